"""Decode / encode rates of the nvJPEG bridge (libpbio.so) on one 8K frame, per backend / decoder:
    PB_IO_BACKEND={default,hybrid,gpu,hardware} PB_IO_DECODER={,gpu,threads} python tests/analysis/jpeg_probe.py [threads]
Smooth synthetic frame (the one bench.py's e2e_compressed uses), quality 90 in / 75 out."""
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    import torch
    from PIL import Image

    from photonbend_b200.utils import image_io

    h, w = 3840, 7680
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx // 8) % 256, (yy // 4) % 256, ((xx + yy) // 16) % 256], axis=2).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=90)
    data = buf.getvalue()
    out = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    tag = os.environ.get("PB_IO_BACKEND", "default") + "/" + (os.environ.get("PB_IO_DECODER") or "gpu")
    try:
        image_io.decode_jpeg_into(data, out)
    except Exception as exc:  # noqa: BLE001
        print(f"{tag:9s} decode refused: {exc}")
        return
    torch.cuda.synchronize()
    ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    err = int(np.abs(out.cpu().numpy().astype(np.int16) - ref).max())
    n = 8
    t0 = time.perf_counter()
    for _ in range(n):
        image_io.decode_jpeg_into(data, out)
    torch.cuda.synchronize()
    dec = (time.perf_counter() - t0) / n
    enc_bytes = image_io.encode_jpeg_from_device(out)
    t0 = time.perf_counter()
    for _ in range(n):
        enc_bytes = image_io.encode_jpeg_from_device(out)
    torch.cuda.synchronize()
    enc = (time.perf_counter() - t0) / n
    px = h * w / 1e9
    print(f"{tag:22s} decode {dec * 1e3:7.2f} ms ({px / dec:5.2f} Gpix/s, max |d| vs Pillow {err}), "
          f"encode {enc * 1e3:7.2f} ms ({px / enc:5.2f} Gpix/s), {len(data)} -> {len(enc_bytes)} bytes")
    n_threads = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    if n_threads > 1:  # several host threads decoding at once into their own tensors, on their own streams
        import threading

        outs = [torch.empty_like(out) for _ in range(n_threads)]

        def work(k):
            with torch.cuda.stream(torch.cuda.Stream()):
                for _ in range(n):
                    image_io.decode_jpeg_into(data, outs[k])

        for warm in (True, False):
            threads = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
            t0 = time.perf_counter()
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        same = all(bool((o == out).all()) for o in outs)
        print(f"{tag:22s} {n_threads} threads: {n_threads * n * px / dt:5.2f} Gpix/s decoded, outputs equal: {same}")


if __name__ == "__main__":
    main()
