"""Image files <-> pixels for the commands: the two spots where the reference touches Pillow
(photonbend/scripts/commands/__init__.py:135-143 ``_open_image`` and the ``Image.fromarray(...)
.save(...)`` at the end of make_pano.py:132-139, alter_photo.py:155-162, make_photo.py:134-141).

Two codecs:

* ``pil`` (default) -- Pillow on the host, exactly what the reference does, so the pixels that
  reach the remap kernel are the reference's pixels.
* ``nvjpeg`` (``PHOTONBEND_B200_CODEC=nvjpeg``) -- JPEG files are decoded straight into a CUDA
  tensor and encoded straight from one by libpbio.so (include/pb_io.h): pixels never visit host
  memory between the file and the kernel.  nvJPEG is not bit-identical to libjpeg-turbo (IDCT
  rounding, chroma upsampling), so outputs differ from the reference's by a few LSB; that is why
  it is opt-in.  PNG files and JPEGs nvJPEG refuses (e.g. CMYK) fall back to Pillow's decoder --
  a codec fall-back on the host side of the file format, not a fall-back of the remap.

Pillow's ``save`` defaults, which the reference uses, are quality 75 and 4:2:0 subsampling; the
nvJPEG encoder is driven with the same.
"""

from __future__ import annotations

import ctypes
import os
import threading
from pathlib import Path

import numpy as np

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG, "libpbio.so")

# every symbol include/pb_io.h declares
EXPORTS = (
    "pb_io_version",
    "pb_io_single_state_decodes",
    "pb_io_last_error",
    "pb_io_jpeg_info",
    "pb_io_jpeg_decode_rgb_u8",
    "pb_io_jpeg_encode_rgb_u8",
)

PB_IO_OK, PB_IO_ERR_INVALID_ARGUMENT, PB_IO_ERR_UNSUPPORTED, PB_IO_ERR_CODEC, PB_IO_ERR_CUDA = range(5)
CSS_444, CSS_422, CSS_420 = 0, 1, 2
PILLOW_DEFAULT_QUALITY = 75

_JPEG_SUFFIXES = (".jpg", ".jpeg")


class CodecError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libpbio error {code}: {message}")
        self.code = code


_lib = None
_encode_buffers = threading.local()


def load_codec():
    """Load (building once if absent) and type libpbio.so.  Loading needs no GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from photonbend_b200 import build as _build

        _build.build_io()
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    lib.pb_io_version.restype = ctypes.c_int
    lib.pb_io_version.argtypes = []
    lib.pb_io_single_state_decodes.restype = ctypes.c_longlong
    lib.pb_io_single_state_decodes.argtypes = []
    lib.pb_io_last_error.restype = ctypes.c_char_p
    lib.pb_io_last_error.argtypes = []
    lib.pb_io_jpeg_info.restype = ctypes.c_int
    lib.pb_io_jpeg_info.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
    lib.pb_io_jpeg_decode_rgb_u8.restype = ctypes.c_int
    lib.pb_io_jpeg_decode_rgb_u8.argtypes = [vp, ctypes.c_size_t, vp, i32, i32, vp]
    lib.pb_io_jpeg_encode_rgb_u8.restype = ctypes.c_int
    lib.pb_io_jpeg_encode_rgb_u8.argtypes = [vp, i32, i32, i32, i32, vp, vp, ctypes.POINTER(ctypes.c_size_t)]
    _lib = lib
    return lib


def _check(lib, code: int) -> None:
    if code != PB_IO_OK:
        raise CodecError(code, (lib.pb_io_last_error() or b"").decode("utf-8", "replace"))


def selected_codec() -> str:
    """'pil' (default) or 'nvjpeg' (PHOTONBEND_B200_CODEC=nvjpeg)."""
    name = os.environ.get("PHOTONBEND_B200_CODEC", "pil").strip().lower()
    if name not in ("pil", "nvjpeg"):
        raise ValueError(f"PHOTONBEND_B200_CODEC must be 'pil' or 'nvjpeg', not {name!r}")
    return name


# ------------------------------------------------------------------------------- nvJPEG


def decode_jpeg_to_device(data: bytes):
    """JPEG bytes -> uint8 CUDA tensor (H, W, 3), RGB.  Raises CodecError when nvJPEG cannot."""
    from photonbend_b200 import engine

    torch = engine._torch()
    lib = load_codec()
    buf = (ctypes.c_ubyte * len(data)).from_buffer_copy(data)
    w, h, n = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _check(lib, lib.pb_io_jpeg_info(buf, len(data), ctypes.byref(w), ctypes.byref(h), ctypes.byref(n)))
    out = torch.empty((h.value, w.value, 3), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    _check(lib, lib.pb_io_jpeg_decode_rgb_u8(buf, len(data), ctypes.c_void_p(out.data_ptr()), w.value, h.value,
                                             ctypes.c_void_p(stream.cuda_stream)))
    stream.synchronize()  # `buf` (the host bitstream) must outlive the decode
    return out


def decode_jpeg_into(data: bytes, out) -> None:
    """JPEG bytes -> the preallocated uint8 CUDA tensor ``out`` (H, W, 3) (a frame of a batch buffer).
    Raises CodecError when nvJPEG cannot, ValueError when the sizes differ."""
    from photonbend_b200 import engine

    torch = engine._torch()
    lib = load_codec()
    buf = (ctypes.c_ubyte * len(data)).from_buffer_copy(data)
    w, h, n = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _check(lib, lib.pb_io_jpeg_info(buf, len(data), ctypes.byref(w), ctypes.byref(h), ctypes.byref(n)))
    if tuple(out.shape) != (h.value, w.value, 3) or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
        raise ValueError(f"decode_jpeg_into: image is {h.value}x{w.value}x3, destination {tuple(out.shape)}")
    with torch.cuda.device(out.device):
        stream = torch.cuda.current_stream()
        _check(lib, lib.pb_io_jpeg_decode_rgb_u8(buf, len(data), ctypes.c_void_p(out.data_ptr()), w.value, h.value,
                                                 ctypes.c_void_p(stream.cuda_stream)))
        stream.synchronize()  # `buf` (the host bitstream) must outlive the decode


def encode_jpeg_from_device(pixels, quality: int = PILLOW_DEFAULT_QUALITY, subsampling: int = CSS_420) -> bytes:
    """uint8 CUDA tensor (H, W, 3) -> baseline JPEG bytes."""
    from photonbend_b200 import engine

    torch = engine._torch()
    if not (engine.is_torch_tensor(pixels) and pixels.is_cuda and pixels.dtype == torch.uint8 and pixels.dim() == 3
            and pixels.shape[2] == 3):
        raise ValueError("encode_jpeg_from_device needs a uint8 CUDA tensor of shape (H, W, 3)")
    pixels = pixels.contiguous()
    lib = load_codec()
    h, w = int(pixels.shape[0]), int(pixels.shape[1])
    # one bitstream buffer per host thread, grown on demand (the library reports the size it needs):
    # a fresh h*w*3-byte buffer per frame is 88 MB of page faults for an 8K image
    capacity = max(getattr(_encode_buffers, "capacity", 0), h * w // 4 + 65536)
    for _ in range(2):
        if getattr(_encode_buffers, "capacity", 0) < capacity:
            _encode_buffers.buffer = (ctypes.c_ubyte * capacity)()
            _encode_buffers.capacity = capacity
        out = _encode_buffers.buffer
        size = ctypes.c_size_t(_encode_buffers.capacity)
        with torch.cuda.device(pixels.device):
            stream = torch.cuda.current_stream()
            code = lib.pb_io_jpeg_encode_rgb_u8(ctypes.c_void_p(pixels.data_ptr()), w, h, int(quality), int(subsampling),
                                                ctypes.c_void_p(stream.cuda_stream), out, ctypes.byref(size))
        if code == PB_IO_ERR_INVALID_ARGUMENT and size.value > _encode_buffers.capacity:
            capacity = int(size.value) + 65536  # too small: the library said how much it needs
            continue
        _check(lib, code)
        return bytes(memoryview(out)[: size.value])
    raise CodecError(PB_IO_ERR_CODEC, "encode_jpeg_from_device: bitstream larger than the size the codec announced")


# ------------------------------------------------------------------------------- files


def open_image(path, codec: str | None = None):
    """Pixels of an image file: a NumPy uint8 array (codec 'pil': whatever Pillow yields, like the
    reference) or a uint8 CUDA tensor (H, W, 3) (codec 'nvjpeg', JPEG files).  IOError when the
    file cannot be read, as Pillow raises it."""
    from PIL import Image

    codec = codec or selected_codec()
    path = Path(path)
    if codec == "nvjpeg" and path.suffix.lower() in _JPEG_SUFFIXES:
        with open(path, "rb") as fh:  # IOError if unreadable, like Image.open
            data = fh.read()
        try:
            return decode_jpeg_to_device(data)
        except CodecError as exc:
            if exc.code not in (PB_IO_ERR_UNSUPPORTED, PB_IO_ERR_CODEC):
                raise
            # a JPEG flavour nvJPEG does not decode: Pillow reads the file instead
    with Image.open(path) as img:
        return np.asarray(img)


def save_image(pixels, path, codec: str | None = None) -> None:
    """Write pixels (NumPy array or CUDA / CPU tensor, uint8 HWC) to a .jpg/.jpeg/.png file.
    IOError when the file cannot be written."""
    from PIL import Image

    from photonbend_b200 import engine

    codec = codec or selected_codec()
    path = Path(path)
    is_tensor = engine.is_torch_tensor(pixels)
    if (codec == "nvjpeg" and path.suffix.lower() in _JPEG_SUFFIXES and is_tensor and pixels.is_cuda
            and pixels.dim() == 3 and pixels.shape[2] == 3):
        data = encode_jpeg_from_device(pixels)
        with open(path, "wb") as fh:
            fh.write(data)
        return
    if is_tensor:
        pixels = pixels.cpu().numpy()
    Image.fromarray(np.ascontiguousarray(pixels)).save(path)


__all__ = ["open_image", "save_image", "decode_jpeg_to_device", "decode_jpeg_into", "encode_jpeg_from_device",
           "selected_codec", "load_codec", "CodecError", "EXPORTS"]
