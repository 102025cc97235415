#!/usr/bin/env python3
"""bench.py -- throughput of the hdr2yuv conversion hot path on B200, beside the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic frames (the default workload is
BASELINE.json configs[1]: a 60-frame 3840x2160 half-float RGB sequence, linear -> PQ, 10-bit
BT.2020 Y'CbCr 4:2:0 with the reference FIR).  Rank 0 prints ONE JSON line:

  value       whole-job Mpixel/s with the batch resident in HBM when the timed region starts
              (statistics + LUT build + fused kernel; CUDA events, max over ranks)
  e2e         the same metric through h2y_forward_host / h2y_inverse_host with HOST buffers
              (pinned): H2D of every source frame and D2H of every .yuv frame inside the timed region
  roofline    the fused kernel alone: algorithmic bytes / its CUDA-event duration,
              against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the reference's own CPU code (oracle/_ref, compiled unmodified) on a bounded sample of
              the same frames on this box's host cores; also used to count code deviations (parity)

Multi-GPU: frames are independent, so each rank converts its own contiguous frame
range; there is no collective on the data path ("scaling": "weak": 60 frames per GPU per step).

`--impl reference` times ONLY the reference's CPU implementation (all host cores, one frame per
core per step) and prints the same line with "impl": "reference".  The oracle is imported here
only for cpu_baseline / --impl reference; the GPU arm runs libhdr2yuv_b200.so and nothing else.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> description of the workload (BASELINE.json configs)
WORKLOADS = {
    # configs[1]: the configuration the metric is quoted on
    "exr4k_pq10_bt2020_420": dict(
        kind="forward", w=3840, h=2160, frames=60, src_kind="half", layout="half_rgb",
        src=dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0),
        dst=dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1),
        text="3840x2160 half RGB (linear, BT.709) -> PQ 10-bit BT.2020nc Y'CbCr 4:2:0 FIR, 60-frame sequence"),
    # configs[4] per-GPU slice
    "exr4k_pq12_bt2020_420": dict(
        kind="forward", w=3840, h=2160, frames=75, src_kind="half", layout="half_rgb",
        src=dict(bit_depth=32, full_range=1, transfer=8, primaries=1, matrix=0),
        dst=dict(bit_depth=12, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1),
        text="3840x2160 half RGB (linear) -> PQ 12-bit BT.2020nc 4:2:0 FIR, 75 frames per GPU (600 over 8)"),
    # configs[0]
    "tiff1080_bt2020_420": dict(
        kind="forward", w=1920, h=1080, frames=120, src_kind="tiff16", layout="rgb16",
        src=dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0),
        dst=dict(bit_depth=10, full_range=0, transfer=16, primaries=9, matrix=9, chroma=1, resampler=1),
        text="1920x1080 16-bit X'Y'Z' PQ TIFF rows -> 10-bit BT.2020nc 4:2:0 FIR, 120 frames"),
    # configs[2]
    "tiff1080_ydzdx_444": dict(
        kind="forward", w=1920, h=1080, frames=120, src_kind="tiff16", layout="rgb16",
        src=dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0),
        dst=dict(bit_depth=10, full_range=0, transfer=16, primaries=10, matrix=11, chroma=3, resampler=1),
        text="1920x1080 16-bit X'Y'Z' TIFF rows -> 10-bit Y'DzDx 4:4:4, 120 frames"),
    "tiff1080_ydzdx_420": dict(
        kind="forward", w=1920, h=1080, frames=120, src_kind="tiff16", layout="rgb16",
        src=dict(bit_depth=16, full_range=0, transfer=16, primaries=10, matrix=0),
        dst=dict(bit_depth=10, full_range=0, transfer=16, primaries=10, matrix=11, chroma=1, resampler=1),
        text="1920x1080 16-bit X'Y'Z' TIFF rows -> 10-bit Y'DzDx 4:2:0 FIR, 120 frames"),
    # configs[3]
    "inverse4k_b10_2020": dict(
        kind="inverse", w=3840, h=2160, frames=60, bit_depth=10, matrix=2, fir=1, full_range=0, alpha=0,
        text="yuv2tiff: 3840x2160 10-bit BT.2020 4:2:0 .yuv -> 4:4:4 FIR upsample -> 16-bit RGB rows, 60 frames"),
    # further yuv2tiff modes (not BASELINE configurations): its default (12-bit Y'DzDx) and the other Y'CbCr depth / family
    "inverse4k_b12_ydzdx": dict(
        kind="inverse", w=3840, h=2160, frames=60, bit_depth=12, matrix=0, fir=1, full_range=0, alpha=0,
        text="yuv2tiff defaults: 3840x2160 12-bit Y'DzDx 4:2:0 .yuv -> 4:4:4 FIR upsample -> 16-bit RGB rows, 60 frames"),
    "inverse4k_b12_2020": dict(
        kind="inverse", w=3840, h=2160, frames=60, bit_depth=12, matrix=2, fir=1, full_range=0, alpha=0,
        text="yuv2tiff B12 2020: 3840x2160 12-bit BT.2020 4:2:0 .yuv -> 4:4:4 FIR upsample -> 16-bit RGB rows, 60 frames"),
    "inverse4k_b10_709": dict(
        kind="inverse", w=3840, h=2160, frames=60, bit_depth=10, matrix=1, fir=1, full_range=0, alpha=0,
        text="yuv2tiff B10 709: 3840x2160 10-bit BT.709 4:2:0 .yuv -> 4:4:4 FIR upsample -> 16-bit RGB rows, 60 frames"),
}
DEFAULT_WORKLOAD = "exr4k_pq10_bt2020_420"
LAYOUT_IDS = {"planar_u16": 0, "planar_f32": 1, "rgb16": 2, "rgba16": 3, "half_rgb": 4, "half_rgba": 5}
LAYOUT_BPP = {"rgb16": 6, "rgba16": 8, "half_rgb": 6, "half_rgba": 8}


# ------------------------------------------------------------------------------ synthetic frames
def make_frame(wl, index):
    """(H,W,C) uint16: what the reader hands to the hot path for frame `index` (seed = index)."""
    from hdr2yuv_b200 import synth
    ch = 4 if wl.get("layout", "").endswith("a") or wl.get("layout") == "rgba16" else 3
    if wl["kind"] == "inverse":
        raise RuntimeError("inverse inputs are produced by the forward path")
    if wl["src_kind"] == "half":
        if wl.get("content") == "flat" and ch == 3:
            return synth.exr_half_frame_flat(wl["w"], wl["h"], seed=index)
        if wl.get("content") in ("natural", "graded") and ch == 3:
            return synth.exr_half_frame_smooth_fast(wl["w"], wl["h"], seed=index, peak_white=wl["content"] == "graded")
        return synth.exr_half_frame_fast(wl["w"], wl["h"], seed=index, channels=ch)
    return synth.tiff16_frame(wl["w"], wl["h"], seed=index + 1, channels=ch)


def fill_frames(wl, first, count, out):
    """out: (count, H, W, C) uint16 view (pinned memory); frames first .. first+count-1."""
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        out[i] = make_frame(wl, first + i)
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        list(ex.map(one, range(count)))


# ------------------------------------------------------------------------------ CPU reference arm
_cpu_state = {}


def _cpu_init():
    from oracle import oracle as O
    _cpu_state["O"] = O
    _cpu_state["kind"] = "reference" if O.ref_available() else "port"
    return _cpu_state["kind"]


def _cpu_warm(_):
    if "O" not in _cpu_state:
        _cpu_init()
        O = _cpu_state["O"]
        O.port_lib()
        if _cpu_state["kind"] == "reference":
            O.ref_lib()
    time.sleep(0.3)          # keeps every pool worker busy so that each one is spawned and warmed
    return os.getpid()


def _cpu_forward_task(args):
    """One frame through the reference's pic_stats -> matrix_convert -> convert -> write_yuv chain.
    Returns (t_start, t_end, yuv or None).  The reader's de-interleave / half->float widening is
    outside the timed region, like file I/O."""
    wl, index, want_output = args
    if "O" not in _cpu_state:
        _cpu_init()
    O = _cpu_state["O"]
    px = make_frame(wl, index)
    planes = O.load_half(px) if wl["src_kind"] == "half" else O.load_rgb16(px, wl["src"]["full_range"])
    backend = "ref" if _cpu_state["kind"] == "reference" else "port"
    t0 = time.monotonic()
    yuv = O.forward(planes, wl["src"], wl["dst"], backend=backend)
    t1 = time.monotonic()
    return t0, t1, (yuv if want_output else None)


def _cpu_inverse_task(args):
    wl, yuv, want_output = args
    if "O" not in _cpu_state:
        _cpu_init()
    O = _cpu_state["O"]
    t0 = time.monotonic()
    rgb, invalid = O.yuv2tiff(yuv, wl["w"], wl["h"], wl["bit_depth"], wl["matrix"], wl["fir"], wl["full_range"],
                              wl["alpha"], backend="port")     # the reference program itself is file-to-file
    t1 = time.monotonic()
    return t0, t1, (rgb if want_output else None)


class CpuPool:
    """All host cores, one single-threaded reference instance per core (the reference has no
    threads; frames are independent, so frame-parallel processes are how a user would run it)."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        self.cores = cores or len(os.sched_getaffinity(0))
        self.ex = ProcessPoolExecutor(max_workers=self.cores, mp_context=mp.get_context("spawn"))
        pids = set(self.ex.map(_cpu_warm, range(self.cores * 2)))
        self.workers = len(pids)
        self.kind = _cpu_init()

    def run(self, fn, tasks):
        """One task per worker, all at once.  Returns (seconds per frame per core, outputs): the mean of the workers' own
        conversion times, so the stagger of the workers' frame synthesis in front of the timed part is not charged to the
        reference."""
        res = list(self.ex.map(fn, tasks))
        wall = sum(r[1] - r[0] for r in res) / len(res)
        return wall, [r[2] for r in res]

    def close(self):
        self.ex.shutdown()


def cpu_sample_tasks(wl, cores, first_frame, want_output, inverse_inputs=None):
    if wl["kind"] == "forward":
        return _cpu_forward_task, [(wl, first_frame + i, want_output) for i in range(cores)]
    return _cpu_inverse_task, [(wl, inverse_inputs[i % len(inverse_inputs)], want_output) for i in range(cores)]


# ------------------------------------------------------------------------------ helpers
def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


def recorded_traffic(name):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # under load = the upper half of the samples (the sampler also sees the idle edges)
        load = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    """torch.distributed is bookkeeping only (barrier + max of the timings); no data-path collective."""
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as d
        torch.cuda.set_device(local)
        d.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = d
    elif n_gpus > 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU); WORLD_SIZE is 1" % n_gpus)
    return dist, rank, local, world


def reduce_max(dist, x, dev):
    if dist is None:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist):
    import torch
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------ the reference arm
def run_reference(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pool = CpuPool()
    inv_inputs = None
    if wl["kind"] == "inverse":
        inv_inputs = [reference_inverse_input(wl, i) for i in range(min(4, pool.cores))]
    px_per_step = wl["w"] * wl["h"] * pool.cores
    total = 0.0
    for s in range(args.warmup + args.steps):
        fn, tasks = cpu_sample_tasks(wl, pool.cores, s * pool.cores, False, inv_inputs)
        wall, _ = pool.run(fn, tasks)
        if s >= args.warmup:
            total += wall
    pool.close()
    mpx = px_per_step * args.steps / total / 1e6
    sample = "%d frames of %dx%d per step (one per host core), %d steps" % (pool.cores, wl["w"], wl["h"], args.steps)
    line = {
        "impl": "reference", "metric": "Mpixel/s", "value": mpx, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64->u16",
        "data": "synthetic", "frames_per_s": mpx * 1e6 / (wl["w"] * wl["h"]),
        "config": bench_config(wl, name), "frames_per_step": pool.cores,
        "cpu_baseline": {"value": mpx, "unit": "Mpixel/s", "cores": pool.cores,
                         "kind": pool.kind if wl["kind"] == "forward" else "port", "sample": sample},
        "e2e": {"value": mpx, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def reference_inverse_input(wl, index):
    """A plausible 10-bit 4:2:0 frame without a GPU: luma gradient + noise, chroma near neutral."""
    rng = np.random.default_rng(100 + index)
    w, h = wl["w"], wl["h"]
    D = 1 << (wl["bit_depth"] - 8)
    y = rng.integers(16 * D, 235 * D + 1, (h, w), dtype=np.uint16)
    c = rng.integers(112 * D, 144 * D + 1, (2, h // 2, w // 2), dtype=np.uint16)
    return np.concatenate([y.reshape(-1), c.reshape(-1)])


def bench_config(wl, name):
    """The workload, identical in both arms (how many frames an arm takes per step is a top-level key of its line)."""
    cfg = {"workload": name, "what": wl["text"], "width": wl["w"], "height": wl["h"],
           "l2": "inputs larger than L2 (batch >> 126 MB); no flush needed"}
    if wl["kind"] == "forward":
        cfg["layout"] = wl["layout"] + " (%d B/px in)" % LAYOUT_BPP[wl["layout"]]
        cfg["content"] = "iid log-uniform samples (SURVEY.md 8d config 2)" if wl["src_kind"] == "half" else "iid uniform codes"
        if wl.get("content") == "natural" and wl["src_kind"] == "half":
            cfg["content"] = "spatially correlated luminance field with colour cast and 2 % noise (--forward-content natural)"
        if wl.get("content") == "flat" and wl["src_kind"] == "half":
            cfg["content"] = "diagnostic: 64x64 tiles of constant colour, broadcast LUT gathers (--forward-content flat)"
        if wl.get("content") == "graded" and wl["src_kind"] == "half":
            cfg["content"] = ("the natural field plus one peak-white and one near-black pixel per frame, so the channels share "
                              "floor and ceiling (--forward-content graded)")
    return cfg


# ------------------------------------------------------------------------------ the GPU arm
def algorithmic_bytes_per_px(wl):
    """every input sample read once, every output sample written once."""
    if wl["kind"] == "inverse":
        return 3 + (8 if wl["alpha"] else 6)
    out = {1: 3, 2: 4, 3: 6}[wl["dst"]["chroma"]]
    return LAYOUT_BPP[wl["layout"]] + out


def fused_kernel_name(wl):
    if wl["kind"] == "inverse":
        return "k_inverse_rows (h2y_inverse.cu)"
    exr = wl["src_kind"] == "half" and wl["dst"]["chroma"] == 1 and wl["dst"]["resampler"] == 1 and \
        wl["src"]["transfer"] != wl["dst"]["transfer"] and wl["dst"]["matrix"] in (1, 9, 11) and wl["dst"]["bit_depth"] <= 12
    return "k_forward_exr420_rows (h2y_forward2.cu)" if exr else "k_forward_fused (h2y_forward.cu)"


def run_gpu(args, wl, name):
    import torch
    from hdr2yuv_b200 import _cabi as cabi
    from hdr2yuv_b200 import api          # raises if libhdr2yuv_b200.so is missing: no fallback

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    dist, rank, local, world = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = api.Context(local)
    if args.plan_reuse != "auto":
        ctx.set_option("H2Y_PLAN_REUSE", args.plan_reuse)
    w, h, nf = wl["w"], wl["h"], min(args.frames or wl["frames"], 256)   # one profiling bracket per step (<= 256 frames)
    px_per_frame = w * h
    from hdr2yuv_b200 import sharding
    first, last = sharding.frame_range(rank, world, nf * world)    # this rank's contiguous frame range (weak scaling)
    assert last - first == nf
    stream = torch.cuda.current_stream()

    # ---- inputs: pinned host batch -> HBM -------------------------------------------------------
    if wl["kind"] == "forward":
        params = api.forward_params(w, h, LAYOUT_IDS[wl["layout"]], wl["src"], wl["dst"],
                                    resampler=wl["dst"]["resampler"],
                                    clip_on_load=1 if wl["src_kind"] == "tiff16" else 0)
        in_bytes = api.src_frame_bytes(params.src)
        out_bytes = api.yuv_frame_bytes(w, h, wl["dst"]["chroma"])
        ch = in_bytes // (px_per_frame * 2)
        h_in = api.PinnedBuffer(in_bytes * nf)
        fill_frames(wl, first, nf, h_in.view(np.uint16).reshape(nf, h, w, ch))
    else:
        # inverse input = forward output of config-2-style frames (realistic chroma, SURVEY.md 8d cfg 4)
        fwl = dict(WORKLOADS["exr4k_pq10_bt2020_420"], w=w, h=h)
        fwl["dst"] = dict(fwl["dst"], bit_depth=wl["bit_depth"], matrix={0: 11, 1: 1, 2: 9, 3: 13, 4: 12}[wl["matrix"]])
        fparams = api.forward_params(w, h, LAYOUT_IDS[fwl["layout"]], fwl["src"], fwl["dst"], resampler=1)
        params = cabi.InverseParams(w, h, wl["bit_depth"], wl["matrix"], wl["fir"], wl["full_range"], wl["alpha"])
        in_bytes = api.yuv_frame_bytes(w, h, cabi.CHROMA_420)
        out_bytes = w * h * (4 if wl["alpha"] else 3) * 2
        h_in = api.PinnedBuffer(in_bytes * nf)
        fb = api.src_frame_bytes(fparams.src)
        tmp_h = np.empty((8, h, w, 3), np.uint16)
        if args.content == "iid":
            fill_frames(fwl, first, 8, tmp_h)
        else:
            from hdr2yuv_b200 import synth
            for i in range(8):
                tmp_h[i] = synth.exr_half_frame_smooth_fast(w, h, seed=first + i)
        d_tmp = torch.from_numpy(tmp_h.view(np.uint8).reshape(-1)).to(dev)
        d_y = torch.empty(in_bytes * 8, dtype=torch.uint8, device=dev)
        ctx.forward(fparams, d_tmp, d_y, 8)
        torch.cuda.synchronize()
        y8 = d_y.cpu().numpy().reshape(8, in_bytes)
        hv = h_in.view(np.uint8).reshape(nf, in_bytes)
        for i in range(nf):
            hv[i] = y8[i % 8]
        del d_tmp, d_y, tmp_h
    h_out = api.PinnedBuffer(out_bytes * nf)
    d_in = torch.empty(in_bytes * nf, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(out_bytes * nf, dtype=torch.uint8, device=dev)
    d_in.copy_(torch.from_numpy(h_in.array), non_blocking=False)
    d_invalid = torch.zeros(nf, dtype=torch.int32, device=dev) if wl["kind"] == "inverse" else None

    def step_device():
        if wl["kind"] == "forward":
            ctx.forward(params, d_in, d_out, nf)
        else:
            ctx.inverse(params, d_in, d_out, nf, invalid=d_invalid)

    def step_host():
        if wl["kind"] == "forward":
            ctx.forward_host(params, h_in.ptr, h_out.ptr, nf)
        else:
            ctx.inverse_host(params, h_in.ptr, h_out.ptr, nf)

    # ---- device-resident: warm-up, then exactly K steps between barriers, CUDA events ---------------
    sampler = ClockSampler(local) if rank == 0 else None       # runs through both timed regions
    for _ in range(args.warmup):
        step_device()
    barrier(dist)
    ctx.profile_enable(True)
    launches0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier(dist)
    dev_ms = reduce_max(dist, e0.elapsed_time(e1), dev)
    launches = ctx.kernel_launches - launches0
    kern_ms, prologue_ms = ctx.profile_last_ms()           # per call, averaged over the last <=16 steps
    ctx.profile_enable(False)
    plan_reuse = None
    if wl["kind"] == "forward":
        attempted, pr_n, pr_redone = ctx.forward_last_plan_reuse()
        plan_reuse = {"single_pass": attempted, "frames": pr_n, "frames_converted_again": pr_redone,
                      "what": "last timed step: frames converted with the previous plan while their extrema were gathered, "
                              "and how many the verify step handed back (h2y_forward, include/hdr2yuv_b200.h)"}
    gpu_ref_out = d_out.cpu().numpy().view(np.uint16).reshape(nf, -1) if rank == 0 and world == 1 else None

    # ---- end to end through the host API: pinned host in, pinned host out --------------------------
    for _ in range(min(args.warmup, 3)):
        step_host()
    barrier(dist)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    barrier(dist)
    e2e_ms = reduce_max(dist, e2e_ms, dev)
    clocks = sampler.stop() if sampler else None
    host_matches_device = None
    if gpu_ref_out is not None:
        host_matches_device = bool(np.array_equal(h_out.view(np.uint16).reshape(nf, -1), gpu_ref_out))

    # ---- what the host platform gives at this N: plain pinned copies, both directions at once, all ranks together ---
    pcie = None
    try:
        nbytes = 1 << 30
        hp_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        hp_out = torch.empty(nbytes // 2, dtype=torch.uint8, pin_memory=True)
        dp_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dp_out = torch.empty(nbytes // 2, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for timed in (False, True):
            barrier(dist)
            t0 = time.perf_counter()
            for _ in range(3):
                with torch.cuda.stream(s1):
                    dp_in.copy_(hp_in, non_blocking=True)
                with torch.cuda.stream(s2):
                    hp_out.copy_(dp_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        dt = reduce_max(dist, dt, dev)
        pcie = {"h2d_gbs_per_gpu": 3 * nbytes / dt / 1e9, "d2h_gbs_per_gpu": 1.5 * nbytes / dt / 1e9,
                "what": "plain cudaMemcpyAsync of pinned buffers, 2:1 H2D:D2H like the workload, all ranks concurrently"}
        del hp_in, hp_out, dp_in, dp_out
    except Exception as exc:                                  # never let the probe break the bench line
        pcie = {"error": str(exc)[:100]}

    # ---- numbers ---------------------------------------------------------------------------------------
    px_step = px_per_frame * nf * world
    mpx = px_step * args.steps / (dev_ms * 1e-3) / 1e6
    e2e_mpx = px_step * args.steps / (e2e_ms * 1e-3) / 1e6
    bpp = algorithmic_bytes_per_px(wl)
    peak, peak_src = measured_hbm_peak()
    achieved = bpp * px_per_frame * nf / (kern_ms * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (the single-pass instantiation
    # has its own entry: it also spills nothing to local memory any more, profiles/r02)
    traffic = recorded_traffic(name + "/single_pass") if plan_reuse and plan_reuse["single_pass"] else None
    if traffic is None:
        traffic = recorded_traffic(name)
    line = {
        "metric": "Mpixel/s", "value": mpx, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/f64->u16", "data": "synthetic",
        "frames_per_s": mpx * 1e6 / px_per_frame,
        "config": bench_config(wl, name), "frames_per_step": nf * world, "plan_reuse": plan_reuse,
        "roofline": {"bound": "hbm", "kernel": fused_kernel_name(wl) + (" <SPEC>: conversion + statistics in one pass" if plan_reuse and plan_reuse["single_pass"] else ""),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": traffic,
                     "algorithmic_bytes_per_px": bpp, "algorithmic_bytes_per_launch": bpp * px_per_frame * nf,
                     "kernel_ms_per_launch": kern_ms, "prologue_ms_per_step": prologue_ms,
                     "kernel_share_of_step": kern_ms / (dev_ms / args.steps),
                     # the whole device-resident step (statistics, plan, LUT and every kernel of the call) against the same
                     # algorithmic bytes: what the single pass is for (0.43 in round 1, two passes over the input)
                     "step_frac": bpp * px_per_frame * nf / (dev_ms / args.steps * 1e-3) / 1e9 / peak,
                     "peak_source": peak_src},
        "e2e": {"value": e2e_mpx, "unit": "Mpixel/s", "frames_per_s": e2e_mpx * 1e6 / px_per_frame,
                "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": in_bytes * nf * world,
                "d2h_bytes_per_step": out_bytes * nf * world,
                "h2d_gbs_per_gpu": in_bytes * nf * args.steps / (e2e_ms * 1e-3) / 1e9,
                "d2h_gbs_per_gpu": out_bytes * nf * args.steps / (e2e_ms * 1e-3) / 1e9,
                "api": "h2y_forward_host" if wl["kind"] == "forward" else "h2y_inverse_host",
                "host_output_equals_device_output": host_matches_device, "pcie_copy_ceiling": pcie,
                # how much of the box's plain-copy ceiling (same ranks, same 2:1 traffic) the API reaches: the end-to-end
                # curve over N is read against the host side of the box (profiles/r02: 52 / 71 / 71 / 91 GB/s aggregate
                # H2D at N = 1 / 2 / 4 / 8 whatever the allocation kind, threads or processes, CPU pinning)
                "frac_of_copy_ceiling": (in_bytes * nf * args.steps / (e2e_ms * 1e-3) / 1e9) / pcie["h2d_gbs_per_gpu"]
                if pcie and pcie.get("h2d_gbs_per_gpu") else None},
        "gpu_launches": int(launches) * world,
        "clocks": clocks,
    }

    # ---- the reference's CPU path on this box's cores: bounded sample of the same frames (N=1 only) --
    if rank == 0 and world == 1 and not args.no_cpu:
        pool = CpuPool()
        n = min(pool.cores, nf)
        if wl["kind"] == "forward":
            fn, tasks = cpu_sample_tasks(wl, n, first, True)
        else:
            hv = h_in.view(np.uint16).reshape(nf, -1)
            fn, tasks = cpu_sample_tasks(wl, n, 0, True, [hv[i].copy() for i in range(n)])
        wall, outs = pool.run(fn, tasks)
        pool.close()
        kind = pool.kind if wl["kind"] == "forward" else "port"    # the reference's yuv2tiff is a file-to-file program
        line["cpu_baseline"] = {
            "value": px_per_frame * n / wall / 1e6, "unit": "Mpixel/s", "cores": n, "kind": kind,
            "sample": "frames 0..%d of this workload (%dx%d), one per host core at once, %.1f s per frame per core" % (n - 1, w, h, wall)}
        differ, max_abs, total = 0, 0, 0
        for i, o in enumerate(outs):
            d = np.abs(gpu_ref_out[i].astype(np.int32) - np.asarray(o).reshape(-1).astype(np.int32))
            differ += int((d != 0).sum()); max_abs = max(max_abs, int(d.max())); total += d.size
        line["parity"] = {"frames": n, "samples": total, "differ": differ, "max_abs_codes": max_abs,
                          "tolerance_codes": 1 if wl["kind"] == "forward" and wl["src"]["transfer"] != wl["dst"]["transfer"] else 0,
                          "against": kind}
    if d_invalid is not None:
        line["config"]["invalid_pixels_per_frame_mean"] = float(d_invalid.float().mean().item())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    h_in.free(); h_out.free()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: the workload's)")
    ap.add_argument("--layout", choices=sorted(LAYOUT_BPP), default=None, help="override the source layout")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--plan-reuse", choices=["auto", "0", "1"], default="auto",
                    help="h2y_forward's single-pass route: library policy (default), never, or whenever a seed exists")
    ap.add_argument("--forward-content", choices=["iid", "natural", "graded", "flat"], default=None,
                    help="forward EXR workloads: iid log-uniform samples (default, the headline; worst case for the LUT "
                         "gather) or a spatially correlated field, as a second data point")
    ap.add_argument("--content", choices=["smooth", "iid"], default="smooth",
                    help="inverse workload only: the .yuv frames come from spatially correlated (default) or iid-random sources")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                                     # timing rule: W >= 3
    wl = dict(WORKLOADS[args.workload])
    if args.layout and wl["kind"] == "forward":
        wl["layout"] = args.layout
    if args.forward_content and wl["kind"] == "forward":
        wl["content"] = args.forward_content
    if args.impl == "reference":
        return run_reference(args, wl, args.workload)
    return run_gpu(args, wl, args.workload)


if __name__ == "__main__":
    sys.exit(main())
