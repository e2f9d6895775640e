#!/usr/bin/env python
"""Where a make-pano call spends its time, per codec (analysis tool; run on the GPU box).

    python tests/analysis/cli_stages.py [image.jpg]

Stages of `make-pano --type inscribed --lens equidistant --fov 360` on the reference's bundled
3072 x 3072 example (tests/golden/equidistant.jpg): decode, remap (upload + kernel + download for
the Pillow codec, kernel only for nvJPEG), encode.  Second of two passes is reported (warm).
"""
from __future__ import annotations

import io
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402


def main():
    import torch

    from photonbend_b200.core import lens
    from photonbend_b200.core.projection import CameraImage, PanoramaImage
    from photonbend_b200.utils import image_io, to_radians

    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "tests", "golden", "equidistant.jpg")
    for codec in ("pil", "nvjpeg"):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pixels = image_io.open_image(path, codec)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            h, w = int(pixels.shape[0]), int(pixels.shape[1])
            src = CameraImage(pixels, to_radians(360), lens.equidistant(), magnitude=w / 2 - 0.5)
            dst = PanoramaImage(np.zeros((h, 2 * h, 3), np.uint8))
            out = src.process_coordinate_map(dst.get_coordinate_map())
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if codec == "nvjpeg":
                data = image_io.encode_jpeg_from_device(out)
            else:
                from PIL import Image

                buf = io.BytesIO()
                Image.fromarray(out).save(buf, format="JPEG")
                data = buf.getvalue()
            t3 = time.perf_counter()
        print(f"{codec:7s} {w}x{h} -> {2 * h}x{h}: decode {1e3 * (t1 - t0):7.1f} ms  remap {1e3 * (t2 - t1):7.1f} ms  "
              f"encode {1e3 * (t3 - t2):7.1f} ms  total {1e3 * (t3 - t0):7.1f} ms  ({len(data) / 1e6:.1f} MB jpeg)")


if __name__ == "__main__":
    main()
