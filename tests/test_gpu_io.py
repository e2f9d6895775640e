"""libpbio.so (include/pb_io.h): JPEG <-> device tensors through nvJPEG.  GPU only.

nvJPEG is a library codec and not bit-identical to libjpeg-turbo, so the checks are tolerances
(stated below), not parity: the parity bar of the remap path is unaffected because Pillow stays
the default codec of the commands."""

import io
import os

import numpy as np
import pytest
from PIL import Image

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def test_device_decode_matches_pillow_within_codec_tolerance():
    """The reference's bundled example (3072 x 3072 JPEG): mean |difference| < 1 level, PSNR > 40 dB
    against Pillow's decode of the same file (IDCT rounding + chroma upsampling differ)."""
    from photonbend_b200.utils import image_io

    path = os.path.join(GOLDEN, "equidistant.jpg")
    with open(path, "rb") as fh:
        dev = image_io.decode_jpeg_to_device(fh.read())
    assert dev.is_cuda and dev.dtype.is_floating_point is False
    got = dev.cpu().numpy()
    want = np.asarray(Image.open(path))
    assert got.shape == want.shape == (3072, 3072, 3)
    assert np.abs(got.astype(np.int16) - want.astype(np.int16)).mean() < 1.0
    assert _psnr(got, want) > 40.0


def test_decode_from_several_host_threads_at_once():
    """Four host threads decode the same file on one GPU, each on its own stream, three times over:
    every output equals the single-threaded one (a decoder state per thread; the single-state path
    holds its state until the stream has run the decode)."""
    import threading

    import torch

    from photonbend_b200.utils import image_io

    with open(os.path.join(GOLDEN, "equidistant.jpg"), "rb") as fh:
        data = fh.read()
    want = image_io.decode_jpeg_to_device(data)
    outs = [torch.zeros_like(want) for _ in range(4)]
    errors = []

    def work(k):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                for _ in range(3):
                    outs[k].zero_()
                    image_io.decode_jpeg_into(data, outs[k])
        except BaseException as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors
    for k in range(4):
        assert torch.equal(outs[k], want), k


def test_device_encode_round_trip():
    """Encode from a CUDA tensor at Pillow's defaults (quality 75, 4:2:0); Pillow decodes the
    bitstream; PSNR against the source > 30 dB on a smooth image, size within 2x of Pillow's own."""
    import torch

    from photonbend_b200.utils import image_io

    yy, xx = np.mgrid[0:480, 0:640]
    img = np.stack([(xx * 255 // 639), (yy * 255 // 479), ((xx + yy) * 255 // 1118)], axis=2).astype(np.uint8)
    data = image_io.encode_jpeg_from_device(torch.from_numpy(img).cuda())
    assert data[:2] == b"\xff\xd8" and data[-2:] == b"\xff\xd9"
    back = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    assert back.shape == img.shape and _psnr(back, img) > 30.0
    ref = io.BytesIO()
    Image.fromarray(img).save(ref, format="JPEG")
    assert 0.5 < len(data) / len(ref.getvalue()) < 2.0


def test_make_pano_with_device_codec(tmp_path, monkeypatch):
    """make-pano with PHOTONBEND_B200_CODEC=nvjpeg: file -> nvJPEG -> remap -> nvJPEG -> file
    without the pixels visiting host memory.  Against the Pillow-codec run of the same command:
    lossless (PNG) outputs differ only by the decoder (mean |difference| < 1.5 levels); the JPEG
    written from the device decodes to within JPEG loss of the PNG (PSNR > 28 dB at quality 75 on
    a 4x nearest-neighbour decimated photo)."""
    from click.testing import CliRunner

    from photonbend_b200.scripts.main import main

    src = os.path.join(GOLDEN, "equidistant.jpg")
    outs = {}
    for codec, suffix in (("pil", "png"), ("nvjpeg", "png"), ("nvjpeg", "jpg")):
        monkeypatch.setenv("PHOTONBEND_B200_CODEC", codec)
        out = tmp_path / f"pano_{codec}.{suffix}"
        res = CliRunner().invoke(main, ["make-pano", "--type", "inscribed", "--lens", "equidistant", "--fov", "360",
                                        "-s", "768", src, str(out)])
        assert res.exit_code == 0, res.output
        outs[codec, suffix] = np.asarray(Image.open(out))
        assert outs[codec, suffix].shape == (768, 1536, 3)
    a, b, c = outs["pil", "png"], outs["nvjpeg", "png"], outs["nvjpeg", "jpg"]
    assert np.abs(a.astype(np.int16) - b.astype(np.int16)).mean() < 1.5
    assert _psnr(c, b) > 28.0
