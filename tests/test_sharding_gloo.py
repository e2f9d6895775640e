"""Multi-GPU path on CPU: two gloo ranks shard a frame stream (k mod N, no collective on the
data path) and reduce their timings with max-over-ranks, exactly as bench.py does under NCCL.
Each rank "remaps" its frames with the oracle so that the union can be checked."""

import os
import socket

import numpy as np
import torch.multiprocessing as mp

from photonbend_b200.batch import shard_frames


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, n_frames, result_dir):
    import torch.distributed as dist

    import case_matrix
    from oracle import numpy_port
    from photonbend_b200.batch import max_over_ranks, shard_frames

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sg = {"kind": "double", "height": 24, "width": 48, "lens": "equidistant", "fov": case_matrix.rad(195)}
    og = {"kind": "equirect", "height": 16, "width": 32}
    mine = list(shard_frames(n_frames, rank, world))
    outs = {k: numpy_port.remap(og, (), sg, case_matrix.case_image(sg, 500 + k)) for k in mine}
    np.savez(os.path.join(result_dir, f"rank{rank}.npz"), **{str(k): v for k, v in outs.items()})
    # timing reduction: every rank must see the slowest rank's time
    slowest = max_over_ranks([10.0 + rank, 5.0 - rank])
    assert slowest == [10.0 + world - 1, 5.0], slowest
    dist.barrier()
    dist.destroy_process_group()


def test_shard_frames_partitions_every_stream():
    for n in (0, 1, 7, 16, 1024):
        for world in (1, 2, 4, 8):
            seen = sorted(k for r in range(world) for k in shard_frames(n, r, world))
            assert seen == list(range(n))
            sizes = [len(shard_frames(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_two_gloo_ranks_cover_the_stream(tmp_path):
    import case_matrix
    from oracle import numpy_port

    world, n_frames = 2, 5
    mp.spawn(_rank_main, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    sg = {"kind": "double", "height": 24, "width": 48, "lens": "equidistant", "fov": case_matrix.rad(195)}
    og = {"kind": "equirect", "height": 16, "width": 32}
    got = {}
    for rank in range(world):
        with np.load(os.path.join(tmp_path, f"rank{rank}.npz")) as z:
            for k in z.files:
                assert int(k) not in got
                got[int(k)] = z[k]
    assert sorted(got) == list(range(n_frames))
    for k in range(n_frames):
        assert np.array_equal(got[k], numpy_port.remap(og, (), sg, case_matrix.case_image(sg, 500 + k)))


def test_shard_rows_partitions_every_frame():
    """Row bands of one frame: disjoint, contiguous, cover [0, H), start on tile rows."""
    from photonbend_b200.batch import ROW_BAND_ALIGN, shard_rows

    for h in (1, 63, 64, 65, 1000, 3840, 4320):
        for world in (1, 2, 3, 4, 8):
            bands = [shard_rows(h, r, world) for r in range(world)]
            rows = [i for b in bands for i in b]
            assert rows == list(range(h)), (h, world)
            assert all(b.start % ROW_BAND_ALIGN == 0 for b in bands if len(b))
    bands = [len(shard_rows(3840, r, 8)) for r in range(8)]
    assert max(bands) - min(bands) <= ROW_BAND_ALIGN


def _band_rank_main(rank, world, port, result_dir):
    import torch.distributed as dist

    import case_matrix
    from oracle import numpy_port
    from photonbend_b200.batch import shard_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sg = {"kind": "camera", "height": 40, "width": 40, "lens": "equidistant", "fov": case_matrix.rad(360), "magnitude": 19.5}
    og = {"kind": "equirect", "height": 150, "width": 64}
    image = case_matrix.case_image(sg, 321)  # every rank holds the whole source
    band = shard_rows(og["height"], rank, world)
    out = numpy_port.remap(og, [(0.2, 0.1, -0.3)], sg, image, rows=(band.start, band.stop))
    np.save(os.path.join(result_dir, f"band{rank}.npy"), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_cover_one_frame_by_row_bands(tmp_path):
    """The other way to shard (north_star: "frames or output-row bands"): two ranks, one frame."""
    import case_matrix
    from oracle import numpy_port

    world = 2
    mp.spawn(_band_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sg = {"kind": "camera", "height": 40, "width": 40, "lens": "equidistant", "fov": case_matrix.rad(360), "magnitude": 19.5}
    og = {"kind": "equirect", "height": 150, "width": 64}
    whole = numpy_port.remap(og, [(0.2, 0.1, -0.3)], sg, case_matrix.case_image(sg, 321))
    got = np.concatenate([np.load(os.path.join(tmp_path, f"band{r}.npy")) for r in range(world)], axis=0)
    assert np.array_equal(got, whole)
