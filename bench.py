#!/usr/bin/env python
"""bench.py -- throughput of the remap path on B200 (and of the reference's CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5|T|cfg1..cfg4]
    python bench.py --impl reference [...]          # the reference's CPU algorithm, host cores
    torchrun ... bench.py --gpus N ...              # one rank per GPU (driver launches this)

One JSON line on stdout (rank 0).  Metric: output Gpix/s, whole job.

A "step" is one pass of the hot path over one batch of ``--frames`` synthetic frames that share
the workload's geometry: ONE pb_remap_u8 launch (the source index of an output pixel is resolved
once and applied to every frame of the batch).  Inputs are resident in HBM before the timed
region; every step touches frames*(src+dst) bytes >> the 126 MB L2, so no L2 flush is needed.
With N ranks every rank owns its own batch (frames sharded k mod N, no collective): weak scaling.

``e2e`` is the same metric through the public API (photonbend_b200.batch.FramePipeline) with
pinned HOST buffers: every frame's H2D copy and D2H copy are inside the timed region.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from photonbend_b200 import workloads  # noqa: E402

METRIC = "output_gpix_per_s"
UNIT = "Gpix/s"
CHANNELS = 3


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------
# accounting


def golden_info(name: str) -> dict:
    key = "cfg4" if name == "cfg5" else name
    with open(os.path.join(REPO, "tests", "golden", "full_configs.json")) as fh:
        return json.load(fh)[key]


def algorithmic_bytes_per_frame(name: str) -> int:
    """B_alg = C*Ho*Wo (every output byte written once) + C*N_touched (every referenced source
    pixel read once); N_touched counted by the oracle (tests/golden/full_configs.json)."""
    info = golden_info(name)
    return CHANNELS * (info["out_pixels"] + info["n_touched"])


def committed_traffic(name: str, frames: int):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/)."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        entry = json.load(fh).get(f"{name}:{frames}")
    return entry["bytes"] if entry else None


def bench_config(name: str, frames: int) -> dict:
    """The workload both arms (--impl b200 / reference) are quoted on."""
    wl = workloads.WORKLOADS[name]
    info = golden_info(name)
    return {"workload": name, "title": wl["title"], "frames_per_step_per_gpu": frames,
            "out": f"{info['shape'][1]}x{info['shape'][0]}x{CHANNELS} u8",
            "l2": "inputs larger than L2: each step reads and writes "
                  f"{frames} distinct frames ({frames * (info['src_pixels'] + info['out_pixels']) * 3 / 1e6:.0f} MB)",
            "sharding": "frames k mod N per GPU, no collective"}


def measured_peak_gbs():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle/numpy_port.py; the live reference
# is pure Python + NumPy and cannot travel to the GPU box)


def _cpu_band_worker(args):
    name, r0, r1 = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import numpy_port

    wl = workloads.WORKLOADS[name]
    image = _CPU_IMAGE[0]
    t0 = time.perf_counter()
    out = numpy_port.remap(wl["out"], wl["rotations"], wl["src"], image, rows=(r0, r1))
    return out.shape[0] * out.shape[1], time.perf_counter() - t0


_CPU_IMAGE = [None]


def cpu_reference_pass(name: str, procs: int, row_fraction: float, repeat: int = 1):
    """One bounded pass of the NumPy port: the first ``row_fraction`` of the output rows of one
    frame, split into ``procs`` row bands run by ``procs`` processes (the reference's protocol
    works on any row band of the map, bit-identically).  Returns (pixels, seconds)."""
    import multiprocessing as mp

    wl = workloads.WORKLOADS[name]
    h = wl["out"]["height"]
    # contiguous bands of >= 64 rows (the per-call set-up of the protocol -- e.g. the flipped copy
    # of the right half of a double image, projection.py:430-431 -- is then a few % of a band),
    # spread evenly over the frame so that cheap (invalid) and expensive rows are both sampled
    band_rows = 64
    n_bands = max(procs, int(round(h * row_fraction / band_rows)))
    n_bands = min(n_bands, h // band_rows)
    stride = h / n_bands
    bands = []
    for k in range(n_bands):
        r0 = min(h - band_rows, int(k * stride))
        bands.append((name, r0, r0 + band_rows))
    bands = bands * max(1, repeat)  # more than one frame's worth of rows: the same frame again
    chunks = 1
    _CPU_IMAGE[0] = workloads.source_image(wl)
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_band_worker, bands, chunksize=chunks)
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt


def cpu_baseline(name: str, budget_s: float = 20.0):
    procs = len(os.sched_getaffinity(0))
    # calibrate on a sliver, then size the sample for ~budget_s of wall time
    px, dt = cpu_reference_pass(name, procs, 0.0)
    rate = px / dt
    h = workloads.WORKLOADS[name]["out"]["height"]
    w = workloads.output_shape(workloads.WORKLOADS[name]["out"])[1]
    frames_worth = rate * budget_s / (h * w)
    frac = min(1.0, max(0.01, frames_worth))
    repeat = max(1, min(64, int(round(frames_worth)))) if frames_worth > 1.0 else 1
    px, dt = cpu_reference_pass(name, procs, frac, repeat)
    return {
        "value": px / dt / 1e9,
        "unit": UNIT,
        "cores": procs,
        "kind": "port",
        "sample": f"oracle/numpy_port.py (NumPy restatement, bit-identical to the reference), "
                  f"{procs} processes over {px // w} output rows ({px / (h * w):.2f} frames of {h} rows) of {name} "
                  f"(64-row bands spread evenly over the frame), {dt:.1f} s wall = {dt * procs:.0f} core-seconds",
    }


def c_port_rate(name: str):
    """Scalar C port of the same algorithm on all cores, for context (not the reference arm)."""
    from oracle import c_port

    wl = workloads.WORKLOADS[name]
    image = workloads.source_image(wl)
    h = wl["out"]["height"]
    rows = (0, h)
    t0 = time.perf_counter()
    out = c_port.remap(wl["out"], wl["rotations"], wl["src"], image, rows=rows, threads=0)
    dt = time.perf_counter() - t0
    return {"value": out.shape[0] * out.shape[1] / dt / 1e9, "unit": UNIT,
            "cores": len(os.sched_getaffinity(0)), "kind": "port",
            "sample": f"oracle/pb_oracle.c, one full {name} frame, {dt:.2f} s wall"}


# --------------------------------------------------------------------------------------------
# clocks


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs.
    NVML is initialised in the constructor (it takes longer than a short timed region), and
    ``start()`` returns only after the first sample is in."""

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self._first = threading.Event()
        self.error = None
        self._nv = self._handle = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self._nv = nv
            self._handle = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._handle, nv.NVML_CLOCK_SM)
            self._names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
        except Exception as exc:  # NVML missing: report it, never fake numbers
            self.error = repr(exc)

    def _sample(self):
        nv, h = self._nv, self._handle
        self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for bit, nm in self._names.items():
            if mask & bit:
                self.reasons.add(nm)

    def run(self):
        if self._nv is None:
            self._first.set()
            return
        try:
            while not self._halt.is_set():
                self._sample()
                self._first.set()
                time.sleep(self.period)
        except Exception as exc:
            self.error = repr(exc)
            self._first.set()

    def start(self):
        super().start()
        self._first.wait(timeout=2)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        out = {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(self.samples)}
        if self.error:
            out["error"] = self.error
        return out


# --------------------------------------------------------------------------------------------
# GPU arm


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            pass
    return local_rank


def make_device_batch(torch, name, frames, rank):
    """``frames`` distinct synthetic frames of the workload, generated on the device from the
    workload's seed (+ global frame index); synthetic uniform noise like the parity inputs."""
    wl = workloads.WORKLOADS[name]
    src = wl["src"]
    gen = torch.Generator(device="cuda")
    gen.manual_seed((wl["seed"] or 1234) + 1000 * rank)
    return torch.randint(0, 256, (frames, src["height"], src["width"], CHANNELS), dtype=torch.uint8,
                         device="cuda", generator=gen)


def timed_kernel_steps(torch, source, cmap, batch, out, steps, warmup, dist, sampler=None):
    from photonbend_b200.batch import remap_batch

    for _ in range(warmup):
        remap_batch(source, cmap, batch, out)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.start()  # clocks are sampled from here to the end of the timed region
    stream = torch.cuda.current_stream()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    from photonbend_b200 import _native

    lib = _native.load()
    launches0 = lib.pb_kernel_launches()
    t_begin.record(stream)
    for k in range(steps):
        starts[k].record(stream)
        remap_batch(source, cmap, batch, out)
        ends[k].record(stream)
    t_end.record(stream)
    torch.cuda.synchronize()
    timed_kernel_steps.launches = lib.pb_kernel_launches() - launches0  # counted by the library itself
    if dist is not None:
        dist.barrier()
    total_ms = t_begin.elapsed_time(t_end)
    launch_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    return total_ms, launch_ms


def timed_e2e_steps(torch, source, cmap, name, frames, steps, warmup, dist):
    """Public-API path with pinned host buffers; H2D + kernel + D2H of every frame timed."""
    from photonbend_b200.batch import FramePipeline

    wl = workloads.WORKLOADS[name]
    src = wl["src"]
    oh, ow, _ = workloads.output_shape(wl["out"])
    pool = min(frames, 4)
    rng = np.random.default_rng(99)
    host_in = []
    for _ in range(pool):
        t = torch.empty((src["height"], src["width"], CHANNELS), dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = rng.integers(0, 256, t.shape, dtype=np.uint8)
        host_in.append(t)
    host_out = [torch.empty((oh, ow, CHANNELS), dtype=torch.uint8, pin_memory=True) for _ in range(pool)]
    pipe = FramePipeline(source, cmap, depth=3)

    def one_step():
        for k in range(frames):
            pipe.submit(host_in[k % pool], host_out[k % pool])
        pipe.drain()

    for _ in range(max(1, min(warmup, 2))):
        one_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    from photonbend_b200 import _native

    lib = _native.load()
    launches0 = lib.pb_kernel_launches()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        dist.barrier()
    h2d = frames * src["height"] * src["width"] * CHANNELS
    d2h = frames * oh * ow * CHANNELS
    return dt, h2d, d2h, lib.pb_kernel_launches() - launches0, host_in[0], host_out[0]


def quick_kernel_rate(torch, name, frames, steps=20, warmup=3):
    """Kernel-only Gpix/s + roofline of another workload (reported under "also")."""
    import helpers
    from photonbend_b200.batch import remap_batch

    wl = workloads.WORKLOADS[name]
    batch = make_device_batch(torch, name, frames, 0)
    source = helpers.product_image(wl["src"], batch)
    cmap = helpers.product_map(wl["out"], wl["rotations"])
    out = remap_batch(source, cmap, batch)
    total_ms, launch_ms = timed_kernel_steps(torch, source, cmap, batch, out, steps, warmup, None)
    info = golden_info(name)
    px = info["out_pixels"] * frames
    avg_ms = float(np.median(launch_ms))
    peak, _ = measured_peak_gbs()
    achieved = algorithmic_bytes_per_frame(name) * frames / (avg_ms * 1e-3) / 1e9
    del batch, out
    torch.cuda.empty_cache()
    return {"title": wl["title"], "frames_per_launch": frames, "value": px / (avg_ms * 1e-3) / 1e9,
            "unit": UNIT, "ms_per_launch": avg_ms, "statistic": "median of 20 launches",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": committed_traffic(name, frames)}}


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload

    cpu = None
    cport = None
    if world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process: the CPU arm forks worker processes
        log(f"[bench] timing the CPU reference arm on {len(os.sched_getaffinity(0))} cores ...")
        cpu = cpu_baseline(name, budget_s=args.cpu_budget)
        cport = c_port_rate(name)

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        # stdout carries ONE JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    import helpers
    from photonbend_b200.batch import remap_batch

    wl = workloads.WORKLOADS[name]
    frames = args.frames
    batch = make_device_batch(torch, name, frames, rank)
    source = helpers.product_image(wl["src"], batch)
    cmap = helpers.product_map(wl["out"], wl["rotations"])
    out = remap_batch(source, cmap, batch)
    torch.cuda.synchronize()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    total_ms, launch_ms = timed_kernel_steps(torch, source, cmap, batch, out, args.steps, args.warmup, dist, sampler)
    clocks = sampler.stop()
    gpu_launches = timed_kernel_steps.launches

    e2e_dt, h2d, d2h, e2e_launches, _, _ = timed_e2e_steps(
        torch, source, cmap, name, frames, max(1, args.e2e_steps), args.warmup, dist)

    from photonbend_b200.batch import max_over_ranks

    total_ms, e2e_ms = max_over_ranks([total_ms, e2e_dt * 1e3], device="cuda")  # timing only

    info = golden_info(name)
    px_per_step = info["out_pixels"] * frames * world
    ms_per_step = total_ms / args.steps
    value = px_per_step / (ms_per_step * 1e-3) / 1e9
    e2e_value = px_per_step * max(1, args.e2e_steps) / (e2e_ms * 1e-3) / 1e9

    peak, peak_src = measured_peak_gbs()
    avg_launch_ms = float(np.mean(launch_ms))
    achieved = algorithmic_bytes_per_frame(name) * frames / (avg_launch_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": args.traffic_bytes if args.traffic_bytes is not None else committed_traffic(name, frames),
        "traffic_source": "ncu --set full capture of one launch, profiles/traffic.json", "peak_source": peak_src,
        "kernel": ("pb::remap_tiled_kernel, %d grid(s) per step (a double-fisheye source is remapped as two grids of "
                   "the same kernel, one per tile class; launch_ms, achieved and traffic cover both)"
                   % max(1, gpu_launches // max(1, args.steps))),
        "launch_ms": avg_launch_ms,
        "algorithmic_bytes_per_launch": algorithmic_bytes_per_frame(name) * frames,
    }

    also = {}
    if world == 1 and not args.no_also:
        for other, fr in (("T", frames), ("T", 1), ("cfg1", 1), ("cfg2", 1), ("cfg3", 1), ("cfg4", 1)):
            if other != name:
                also[other if (fr == 1) == (other != "T") else f"{other}_{fr}frame"] = quick_kernel_rate(torch, other, fr)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "dtype_note": "float64 index arithmetic on uint8 pixels",
            "data": "synthetic uniform-noise uint8 frames (seeded), generated on device",
            "config": bench_config(name, frames),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "photonbend_b200.batch.FramePipeline (pinned host frames in and out, depth 3)",
                    "steps": max(1, args.e2e_steps)},
            "gpu_launches": gpu_launches,
            "e2e_gpu_launches": e2e_launches,
            "clocks": clocks,
            "roofline": roofline,
        }
        if cpu is not None:
            cpu["c_port_all_cores"] = cport
            line["cpu_baseline"] = cpu
        if also:
            line["also"] = also
        emit_line(line)
    if dist is not None:
        dist.destroy_process_group()


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (NumPy port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    procs = len(os.sched_getaffinity(0))
    wl = workloads.WORKLOADS[name]
    # size one step for roughly cpu_budget / (steps + warmup) seconds
    px, dt = cpu_reference_pass(name, procs, 0.0)
    rate = px / dt
    h = wl["out"]["height"]
    w = workloads.output_shape(wl["out"])[1]
    per_step_s = max(1.0, args.cpu_budget * 6 / (args.steps + args.warmup))
    frames_worth = rate * per_step_s / (h * w)
    frac = min(1.0, max(0.005, frames_worth))
    repeat = max(1, min(64, int(round(frames_worth)))) if frames_worth > 1.0 else 1
    for _ in range(args.warmup):
        cpu_reference_pass(name, procs, frac, repeat)
    tot_px, tot_dt = 0, 0.0
    for _ in range(args.steps):
        px, dt = cpu_reference_pass(name, procs, frac, repeat)
        tot_px += px
        tot_dt += dt
    value = tot_px / tot_dt / 1e9
    sample = (f"oracle/numpy_port.py (NumPy restatement of the reference, bit-identical), {procs} processes, "
              f"each step = {px // w} output rows ({px / (h * w):.2f} frames of {h} rows) of {name} "
              f"(64-row bands spread evenly), {tot_dt / args.steps:.1f} s wall per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "dtype_note": "float64 index arithmetic on uint8 pixels",
        "data": "synthetic uniform-noise uint8 frame (seeded)",
        "config": bench_config(name, args.frames),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


_RESULT_FD = None


def claim_stdout():
    """stdout carries ONE JSON line: everything else that writes to fd 1 (NCCL's version banner,
    library chatter of child processes) goes to stderr from here on."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(workloads.WORKLOADS))
    ap.add_argument("--frames", type=int, default=16, help="frames per step per GPU (one launch)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU-arm wall time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
