#!/usr/bin/env python
"""A handful of small remaps that touch every kernel, for compute-sanitizer (run on the GPU box):

    compute-sanitizer --tool memcheck  python tests/analysis/sanitize_cases.py
    compute-sanitizer --tool racecheck python tests/analysis/sanitize_cases.py
    compute-sanitizer --tool synccheck python tests/analysis/sanitize_cases.py
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    import torch

    import case_matrix
    import helpers
    from oracle import numpy_port
    from photonbend_b200.batch import remap_row_band

    rad = case_matrix.rad
    cam = {"kind": "camera", "height": 208, "width": 208, "lens": "equidistant", "fov": rad(360), "magnitude": 103.5}
    dbl = {"kind": "double", "height": 176, "width": 352, "lens": "equidistant", "fov": rad(195)}
    pano = {"kind": "equirect", "height": 128, "width": 256}
    eq_out = {"kind": "equirect", "height": 200, "width": 272}   # 4 x 9 tiles with partial edges
    cam_out = {"kind": "camera", "height": 136, "width": 144, "lens": "equisolid", "fov": rad(180), "magnitude": 67.5}
    rot = [(0.3, -0.7, 1.1)]
    checks = 0
    for sg in (cam, dbl, pano):
        frames = np.stack([case_matrix.case_image(sg, 40 + k) for k in range(3)])
        dev = torch.from_numpy(frames).cuda()
        for og, rots in ((eq_out, ()), (eq_out, rot), (cam_out, rot)):
            cmap = helpers.product_map(og, rots)
            single = helpers.product_image(sg, dev[0]).process_coordinate_map(cmap)      # sep1 / direct
            batch = helpers.product_image(sg, dev).process_coordinate_map(cmap)          # tiled (lean / general)
            band = remap_row_band(helpers.product_image(sg, dev[0]), cmap, dev[0], range(64, og["height"]))
            torch.cuda.synchronize()
            want = numpy_port.remap(og, rots, sg, frames[0])
            bad = int((single.cpu().numpy() != want).any(axis=2).sum())
            assert bad <= 3, (sg["kind"], og["kind"], len(rots), bad)
            assert np.array_equal(batch[0].cpu().numpy(), single.cpu().numpy())
            assert np.array_equal(band.cpu().numpy(), single.cpu().numpy()[64:])
            checks += 1
    grey = case_matrix.case_image(cam, 1, 1)
    helpers.product_remap(eq_out, rot, cam, grey)                                         # generic kernel, C = 1
    np.asarray(helpers.product_map(cam_out, rot))                                          # materialised map
    torch.cuda.synchronize()
    print(f"{checks} geometry combinations through every kernel: ok")


if __name__ == "__main__":
    main()
