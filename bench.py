#!/usr/bin/env python
"""Benchmark of the Phase 2 -> 3 hot path (BASELINE.json): frames/sec of DETR-ResNet-50 at 800x1333, bf16, batch 64
per GPU, followed by foot-point homography + 16-zone classification + per-timestamp counting.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's CPU path on the host cores)
    python bench.py --config {1,3,4,5} ...                   (BASELINE.json configs; default 3 = the headline)

Configs (BASELINE.json `configs`):
  3 (default)  configs[1] + configs[2]: DETR-R50 bf16 batch 64 at 800x1333 per GPU + homography + 16 zones + counts [64,17];
               one "step" = one pass over one batch of 64 synthetic frames per rank (weak scaling), all-reduce of the histogram.
               The default line also carries the other GPU configs as `extra.config1 / config4 / config5` (device-resident
               numbers, bounded in time), so that one driver run sees all of them; --no-extras drops them.
  1            1 synthetic 1280x720 frame + homography + 4-zone count, batch 1 (the reference's own CPU-runnable case).
  4            frame-sharded timelapse run: 12 500 frames per GPU (100 000 at 8 GPUs) generated ON THE DEVICE, one timestamp
               slot per frame, hist [12 500 N, 17] per rank, ONE all-reduce at the end; frames/s over the whole run.
  5            homography + point-in-polygon + zone-count microbench: 10^8 synthetic points x 64 polygons, N/G points per rank
               (strong scaling) + all-reduce of [1,65]; points/s and the fraction of the measured HBM peak.
Prints ONE JSON line (rank 0).  See DESIGN.md "measurement" for what each field means.
"""

from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "frames/sec DETR-R50 800x1333 bf16"
UNIT = "frames/s"
METRIC5 = "points/sec homography + point-in-polygon zone count, 64 polygons"
UNIT5 = "points/s"
GFLOP_PER_FRAME = 203.18          # algorithmic, 2 flops per MAC, convs + linears + attention matmuls
GFLOP_PER_FRAME_720P = 191.96     # 1280x720 camera frame -> 3x750x1333
H_IN, W_IN, BATCH, N_ZONES = 800, 1333, 64, 16
CONFIG4_FRAMES_PER_GPU = 12_500   # 100 000 frames over 8 GPUs
CONFIG4_SEED_BASE = 1000
CONFIG5_POINTS = 100_000_000
CONFIG5_ZONES = 64
BYTES_PER_POINT = 12              # 8 B foot point read + 4 B zone index written


def peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def finish(self) -> dict:
        self.stop_flag.set()
        self.join(timeout=2)
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 6 and s[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU implementation of the path on the host cores
# --------------------------------------------------------------------------------------------------------------
def cpu_detect_run(steps: int, warmup: int, frames_per_step: int, h: int, w: int, n_zones: int, keep_first: bool = False) -> dict:
    """Phase 2 -> 3 on the CPU.  Detection leg: transformers' DetrImageProcessor + DetrForObjectDetection in float32 on all host
    threads - the third-party arithmetic the reference's removed ViTDetector drove (kind "reference": the installed library
    itself, random-init weights, synthetic frames).  Floor leg: the reference's transform_batch + classify + get_zone_counts as
    restated by oracle/floor_oracle.c (kind "port": the reference's own sources are not on the GPU box; the C port is FASTER than
    the reference's Python loops, so the baseline is conservative).  Returns frames/s over `steps` steps."""
    import numpy as np
    import torch

    from office_person_detection_vit_b200.detection.synthetic import random_init_state_dict, synthetic_frames
    from office_person_detection_vit_b200.scene import H_CONFIG, grid_zones
    from oracle import detr_oracle as do
    from oracle import floor_oracle as fo

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = do.hf_model(random_init_state_dict(0))
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    from transformers import DetrImageProcessor

    proc = DetrImageProcessor()
    zones = grid_zones(n_zones)
    frames = synthetic_frames(frames_per_step, h, w, seed=1)
    first = {}

    def one_step():
        rgb = [np.ascontiguousarray(f[:, :, ::-1]) for f in frames]
        with torch.no_grad():
            inp = proc(images=rgb, return_tensors="pt")
            out = model(**inp)
            res = proc.post_process_object_detection(out, threshold=0.5, target_sizes=torch.tensor([[h, w]] * len(rgb)))
        if keep_first and not first:
            first["logits"], first["boxes"] = out.logits.clone(), out.pred_boxes.clone()
        n = 0
        for r in res:
            keep = r["labels"] == do.PERSON_LABEL
            b = r["boxes"][keep].double().numpy()
            if len(b):
                xywh = np.stack([b[:, 0], b[:, 1], b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], axis=1)
                px, _, _ = fo.transform(H_CONFIG, xywh, is_bbox=True)
                idx, _ = fo.classify(px, zones)
                fo.count(zone_idx=idx, Z=n_zones)
                n += len(b)
        return n

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    return {"value": steps * frames_per_step / dt, "ms_per_step": dt / steps * 1e3, "cores": cores, "warmup_run": warmup,
            "first": first,
            "kind": "reference", "legs": {"detection": "reference (installed transformers: DetrImageProcessor + DetrForObjectDetection, "
                                                       "fp32 eager, all host threads)",
                                          "floor": "port (oracle/floor_oracle.c restatement of transform_batch / classify / "
                                                   "get_zone_counts, single thread)"},
            "sample": f"{steps} step(s) x {frames_per_step} synthetic {h}x{w} frame(s) after {warmup} warm-up step(s): transformers "
                      f"DetrImageProcessor + DetrForObjectDetection fp32 eager on {cores} threads + oracle port of transform_batch / "
                      f"classify / get_zone_counts ({n_zones} zones)"}


def cpu_floor_run(n_points: int, n_zones: int, py_points: int = 20_000) -> dict:
    """Config 5 on the CPU: the floor path on a bounded sample of the same point distribution.  `value` = the C port
    (oracle/floor_oracle.c, single thread); `python_loop_value` = the reference-shaped pure-Python loop (the reference's
    ZoneClassifier.classify IS a Python loop over zones x edges) on a smaller sample."""
    import numpy as np

    from office_person_detection_vit_b200.scene import H_CONFIG, camera_points, grid_zones
    from oracle import floor_oracle as fo

    zones = grid_zones(n_zones)
    pts = camera_points(n_points, seed=3)
    fo.project_classify_count(H_CONFIG, pts[:1000], zones)
    t0 = time.perf_counter()
    fo.project_classify_count(H_CONFIG, pts, zones)
    dt = time.perf_counter() - t0
    small = pts[:py_points].astype(np.float64)
    t1 = time.perf_counter()
    px = fo.transform_np(H_CONFIG, small, is_bbox=False)
    polys = [z["polygon"] for z in zones]
    counts = [0] * (n_zones + 1)
    for x, y in px:
        hit = [i for i, poly in enumerate(polys) if fo.point_in_polygon_py(float(x), float(y), poly)]
        counts[min(hit, key=lambda i: (zones[i]["priority"], i)) if hit else n_zones] += 1
    dpy = time.perf_counter() - t1
    return {"value": n_points / dt, "unit": UNIT5, "cores": 1, "kind": "port",
            "python_loop_value": py_points / dpy,
            "sample": f"first {n_points} of the synthetic camera points through oracle/floor_oracle.c (C port of transform_batch + "
                      f"classify + get_zone_counts, {n_zones} grid zones, single thread); python_loop_value: the reference-shaped "
                      f"pure-Python loop on {py_points} points"}


def main_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = max(1, args.warmup)
    if args.config == 5:
        r = cpu_floor_run(1_000_000, CONFIG5_ZONES)
        line = {"impl": "reference", "metric": METRIC5, "value": r["value"], "unit": UNIT5, "n_gpus": args.gpus, "steps": 1,
                "warmup": 1, "ms_per_step": 1e9 / r["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config5_workload(args.gpus, CONFIG5_POINTS),
                "cpu_baseline": r, "e2e": {"value": r["value"], "unit": UNIT5, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return
    if args.config == 1:
        r = cpu_detect_run(max(1, args.steps), warm, 1, 720, 1280, 4)
        cfg = config1_workload()
    else:
        # batch 8 on the CPU (the GPU arm's batch of 64 is the same workload, 8 frames are a bounded sample of it)
        r = cpu_detect_run(max(1, args.steps), warm, 8, H_IN, W_IN, N_ZONES)
        cfg = workload_config(args.gpus) | {"reference_frames_per_step": 8}
    r.pop("first", None)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "legs": r["legs"],
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(n_gpus: int) -> dict:
    return {"workload": "BASELINE configs[1] DETR-ResNet-50 (random-init) bf16 batch=64 at 800x1333 detection per GPU, "
                        "followed by the configs[2] tail (foot-point homography + 16-polygon zone classification + "
                        "per-frame counts [64,17]); frames sharded across GPUs, one NCCL all-reduce of the histogram per step",
            "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "frame": [H_IN, W_IN, 3], "zones": N_ZONES,
            "cache": "inputs larger than L2 (205 MB of frames, GBs of activations per step)",
            "parallelism": f"frames dp{n_gpus}"}


def config1_workload() -> dict:
    return {"workload": "BASELINE configs[0]: DETR-ResNet-50 (random-init) person detection batch=1 on one synthetic 1280x720 frame "
                        "(-> 3x750x1333) + homography + 4-zone count", "batch_per_gpu": 1, "global_batch": 1,
            "frame": [720, 1280, 3], "zones": 4, "cache": "L2 flushed between timed steps", "parallelism": "single frame"}


def config4_workload(n_gpus: int, frames_per_gpu: int) -> dict:
    return {"workload": f"BASELINE configs[3]: frame-sharded timelapse run, {frames_per_gpu} synthetic 800x1333 frames per GPU "
                        f"({frames_per_gpu * n_gpus} in total; 100 000 at 8 GPUs) generated on the device from seed = 1000 + "
                        "global_frame // 64, batches of 64, detect + homography + 16-zone classification, one timestamp slot per "
                        f"frame, hist [{frames_per_gpu * n_gpus}, 17] int32 per rank, ONE NCCL all-reduce at the end",
            "frames_per_gpu": frames_per_gpu, "frames_total": frames_per_gpu * n_gpus, "batch_per_gpu": BATCH, "zones": N_ZONES,
            "frame": [H_IN, W_IN, 3], "cache": "every batch is new data (205 MB generated per step)", "parallelism": f"frames dp{n_gpus}"}


def config5_workload(n_gpus: int, n_points: int) -> dict:
    return {"workload": f"BASELINE configs[4]: homography + point-in-polygon zone-count microbench, {n_points} synthetic camera points "
                        f"(x~U[0,1280), y~U[0,720), float32, device Philox generator seed 3) x {CONFIG5_ZONES} polygons (grid_zones), "
                        f"allow_overlap=False, single timestamp slot, {n_points // n_gpus} points per GPU + all-reduce of [1,65]",
            "points": n_points, "points_per_gpu": n_points // n_gpus, "zones": CONFIG5_ZONES,
            "bytes_per_point": BYTES_PER_POINT, "cache": f"inputs larger than L2 ({n_points // n_gpus * 8 / 1e6:.0f} MB of points per GPU)",
            "parallelism": f"points dp{n_gpus}"}


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
class Ctx:
    """Rank / device / collectives of one bench process."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU path)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            # NCCL prints its version banner on stdout when NCCL_DEBUG is set on the box: keep stdout for the one JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = torch.device("cuda", self.local)
        self._flush = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def flush_l2(self):
        """Write a buffer larger than the 126 MB L2 (for the configs whose inputs are smaller than L2)."""
        if self._flush is None:
            self._flush = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
        self._flush.add_(1)

    def timed(self, fn, steps: int) -> float:
        """Total device time (ms, max over ranks) of `steps` calls of fn, CUDA events on the launching stream, bracketed by a
        barrier + synchronize on both sides."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def close(self):
        """End of the run.  Multi-GPU: the captured CUDA graphs hold NCCL kernels and the ranks finish at different times (rank 0
        still profiles and runs the CPU legs after the others are done); tearing the communicator down rank by rank under those
        conditions hung on the 2-GPU box (r02), so every rank flushes, synchronises its device and leaves without the collective
        teardown - the driver reclaims the context."""
        self.torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        if self.world > 1:
            os._exit(0)


def build_pipeline(ctx: Ctx, n_zones: int, batch: int):
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ViTDetector
    from office_person_detection_vit_b200.detection.synthetic import random_init_state_dict
    from office_person_detection_vit_b200.pipeline import DetectCountPipeline
    from office_person_detection_vit_b200.scene import H_CONFIG, grid_zones
    from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer
    from office_person_detection_vit_b200.zone import ZoneClassifier

    # development A/B switch: OPD_OPTIONS="name=value,..." -> opd_set_option before the plans are built; recorded in `config`
    options = dict(kv.split("=") for kv in os.environ.get("OPD_OPTIONS", "").split(",") if kv)
    for k, v in options.items():
        _lib.check(_lib.lib().opd_set_option(k.encode(), int(v)), f"opd_set_option({k})")
    det = ViTDetector(confidence_threshold=0.5, state_dict=random_init_state_dict(0), device=f"cuda:{ctx.local}", batch_size=batch)
    det.load_model()
    pipe = DetectCountPipeline(det, HomographyTransformer(H_CONFIG, FloorMapConfig()),
                               ZoneClassifier(grid_zones(n_zones), allow_overlap=False))
    return det, pipe, options


_GROUPS = [(re.compile(r"^stage(\d)\.\d+\.tail"), lambda m: f"stage{int(m.group(1)) + 1}.tail(3x3+1x1b fused)"),
           (re.compile(r"^stage(\d)\.\d+\.(conv\w+|shortcut)"), lambda m: f"stage{int(m.group(1)) + 1}.{m.group(2)}"),
           (re.compile(r"^enc\d\.attn"), lambda m: "encoder.attention"),
           (re.compile(r"^enc\d\."), lambda m: "encoder.gemm"),
           (re.compile(r"^dec"), lambda m: "decoder")]


def per_kernel_roofline(steps: list[dict], pk: dict) -> dict:
    """Per layer group of one forward (CUDA events between launches): time, algorithmic TFLOP/s and GB/s, and both roofline
    fractions (sustained bf16 peak / measured HBM copy peak); `bound` names the larger one."""
    groups: dict[str, dict] = {}
    for st in steps:
        name = st["name"]
        for rx, fn in _GROUPS:
            m = rx.match(name)
            if m:
                name = fn(m)
                break
        g = groups.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        g["launches"] += 1
        g["ms"] += st["ms"]
        g["flops"] += st["flops"]
        g["bytes"] += st["bytes"]
    out = {}
    for name, g in groups.items():
        tf = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        gb = g["bytes"] / (g["ms"] * 1e-3) / 1e9 if g["ms"] > 0 else 0.0
        ft, fh = tf / pk["bf16_tflops_sustained"], gb / pk["hbm_gbs"]
        out[name] = {"launches": g["launches"], "ms": round(g["ms"], 4), "tflops": round(tf, 1), "frac_tensor": round(ft, 3),
                     "gbs": round(gb, 1), "frac_hbm": round(fh, 3), "bound": "tensor" if ft >= fh else "hbm"}
    return out


def parity_block(out: dict, frames_host, n_frames: int = 1) -> dict:
    """The bench batch's first frame(s) against the CPU oracle (the checker, never the thing measured): oracle "bf16" mode = the
    CUDA path's rounding points, "fp32" = the reference arithmetic (pinned to transformers' DetrForObjectDetection)."""
    from office_person_detection_vit_b200.detection.synthetic import random_init_state_dict
    from oracle import detr_oracle as do

    w = random_init_state_dict(0)
    fr = frames_host[:n_frames].numpy()
    logits, boxes = out["logits"][:n_frames].float().cpu(), out["boxes"][:n_frames].float().cpu()
    sc, lb, xy = do.postprocess(logits, boxes, H_IN, W_IN)
    block = {"config": f"frame(s) 0..{n_frames - 1} of the timed batch ({H_IN}x{W_IN}), all 100 queries before thresholding, seeded "
                       "random-init weights (high gain: see DESIGN.md numerics)",
             "contract": "north_star: boxes <= 1e-2 px, scores <= 1e-3 (not reachable with bf16 storage: DESIGN.md numerics states "
                         "the bound bf16 admits)"}
    for mode in ("bf16", "fp32"):
        rl, rb = do.forward(w, fr, mode=mode)
        rs, rlab, rx = do.postprocess(rl, rb, H_IN, W_IN)
        e = (xy - rx).abs()
        block[f"vs_oracle_{mode}"] = {"box_px_max": round(float(e.max()), 4), "box_px_median": round(float(e.median()), 4),
                                      "score_max": round(float((sc - rs).abs().max()), 5),
                                      "score_median": round(float((sc - rs).abs().median()), 5),
                                      "label_agreement": round(float((lb == rlab).float().mean()), 4)}
    return block


def main_headline(args) -> None:
    ctx = Ctx()
    torch = ctx.torch
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection.synthetic import synthetic_frames

    det, pipe, options = build_pipeline(ctx, N_ZONES, BATCH)
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    B = args.batch
    host = torch.from_numpy(synthetic_frames(B, H_IN, W_IN, seed=1 + rank)).pin_memory()
    frames = host.to(dev)
    hist = torch.zeros(B * world, N_ZONES + 1, dtype=torch.int32, device=dev)

    def step(fr):
        hist.zero_()
        out = pipe.run_tensors(fr, hist=hist, slot_base=rank * B)
        pipe.all_reduce(hist)
        return out

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(frames)
    ctx.barrier()

    # ---- device-resident throughput ----
    # The step's launches (all on one stream, tensor maps pre-encoded, no host synchronisation) are captured ONCE into a CUDA
    # graph and replayed per step.  Multi-GPU: the NCCL all-reduce of the histogram is captured as the graph's last node; if that
    # capture fails the graph holds the kernels only and the all-reduce follows each replay.  --no-graph times plain launches.
    launches0 = _lib.lib().opd_launch_count()
    out = step(frames)
    launches_per_step = _lib.lib().opd_launch_count() - launches0
    graph, graph_has_allreduce = None, False
    if not args.no_graph:
        for with_ar in ([True, False] if world > 1 and not args.no_graph_allreduce else [False]):
            try:
                graph = pipe.capture(frames, hist=hist, slot_base=rank * B, zero_hist=True, all_reduce=with_ar)
                graph_has_allreduce = with_ar
                break
            except Exception as e:   # capture is an optimisation, never a requirement
                print(f"bench.py: CUDA graph capture (all_reduce={with_ar}) failed ({type(e).__name__}: {e})", file=sys.stderr)
                graph = None
                torch.cuda.synchronize()
    if world > 1:   # every rank must take the same path
        flag = torch.tensor([int(graph is not None), int(graph_has_allreduce)], device=dev)
        ctx.dist.all_reduce(flag, op=ctx.dist.ReduceOp.MIN)
        ok_graph, ok_ar = bool(flag[0].item()), bool(flag[1].item())
        if graph is not None and graph_has_allreduce and not ok_ar:
            graph = pipe.capture(frames, hist=hist, slot_base=rank * B, zero_hist=True) if ok_graph else None
            graph_has_allreduce = False
        if not ok_graph:
            graph = None

    def timed_step():
        if graph is None:
            return step(frames)
        o = graph()
        if not graph_has_allreduce:
            pipe.all_reduce(hist)
        return o

    for _ in range(2):
        out = timed_step()
    sampler = ClockSampler(ctx.local)
    sampler.start()
    total_ms = ctx.timed(timed_step, args.steps)
    clocks = sampler.finish()
    launches = launches_per_step * args.steps
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms / 1e3)
    n_det = int(out["n_keep"].sum().item())

    # ---- end to end through the public tensor API with HOST buffers (pinned): double-buffered H2D on a copy stream, the step,
    # and a D2H read of the step's RESULT - the compacted per-frame detections (boxes, scores, foot points, counts) and the
    # per-frame zone counts - all inside the timed region ----
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [torch.empty_like(frames), torch.empty_like(frames)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    result_keys = ("det_xywh", "det_score", "det_foot", "n_keep", "zone_idx")
    host_out = [{k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in result_keys} |
                {"hist": torch.empty(B, N_ZONES + 1, dtype=torch.int32).pin_memory()} for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            dev_bufs[i % 2].copy_(host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    e2e_graphs = [None, None]
    if graph is not None:
        try:
            e2e_graphs = [pipe.capture(dev_bufs[k], hist=hist, slot_base=rank * B, zero_hist=True, all_reduce=graph_has_allreduce)
                          for k in range(2)]
        except Exception as e:
            print(f"bench.py: CUDA graph capture (e2e) failed ({type(e).__name__}: {e}); plain launches", file=sys.stderr)
            e2e_graphs = [None, None]
            torch.cuda.synchronize()

    def e2e_step(k):
        if e2e_graphs[k] is None:
            return step(dev_bufs[k])
        o = e2e_graphs[k]()
        if not graph_has_allreduce:
            pipe.all_reduce(hist)
        return o

    def e2e_run(n):
        for c in consumed:
            c.record()
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            o = e2e_step(i % 2)
            consumed[i % 2].record()
            ho = host_out[i % 2]
            for k in result_keys:
                ho[k].copy_(o[k], non_blocking=True)
            ho["hist"].copy_(hist[rank * B:(rank + 1) * B], non_blocking=True)
        torch.cuda.synchronize()

    e2e_run(2)
    ctx.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    e2e_run(args.steps)
    t1.record()
    ctx.barrier()
    wall = time.perf_counter() - w0
    e2e_ms = ctx.max_over_ranks(max(t0.elapsed_time(t1), wall * 1e3))
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    h2d = host.numel()
    d2h = sum(t.numel() * t.element_size() for t in host_out[0].values())

    # ---- end to end through the reference-shaped OBJECT API: detect_batch(list[ndarray]) -> list[list[Detection]] ----
    frame_list = [f for f in host.numpy()]
    det.detect_batch(frame_list[:B])
    ctx.barrier()
    n_obj = 3
    w0 = time.perf_counter()
    for _ in range(n_obj):
        dets = det.detect_batch(frame_list)
    torch.cuda.synchronize()
    obj_s = ctx.max_over_ranks((time.perf_counter() - w0) * 1e3) / 1e3
    e2e_objects = {"value": round(world * B * n_obj / obj_s, 2), "unit": UNIT,
                   "api": "ViTDetector.detect_batch(list of 64 uint8 ndarrays) -> list[list[Detection]] (pinned staging, H2D, forward, "
                          "post-processing, D2H, one Python Detection object per kept box), wall clock",
                   "detections_per_call": sum(len(d) for d in dets)}

    extra = {}
    if not args.no_extras:
        for name, fn in (("config5", lambda: run_config5(ctx, CONFIG5_POINTS, max(5, min(args.steps, 20)), with_e2e=False)),
                         ("config4", lambda: run_config4(ctx, det, pipe, args.config4_frames, graph=not args.no_graph)),
                         ("config1", lambda: run_config1(ctx, det, steps=max(5, min(args.steps, 20))))):
            try:
                extra[name] = fn()
            except Exception as e:   # an extra never takes the headline down with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"}
                torch.cuda.synchronize()

    if rank != 0:
        ctx.close()
        return

    # ---- roofline of the dominant kernels (the tcgen05 GEMM / convolution family), CUDA events between launches ----
    pk, pk_kind = peaks()
    pipe.run_tensors(frames, hist=hist, slot_base=rank * B)   # local only: the other ranks have left, no collective here
    torch.cuda.synchronize()
    profs = [det.model.profile() for _ in range(3)]
    n_steps = len(profs[0])
    med = [statistics.median(p[i]["ms"] for p in profs) for i in range(n_steps)]
    per_launch = [{**st, "ms": med[i]} for i, st in enumerate(profs[0])]
    by_kind: dict[str, dict] = {}
    for st in per_launch:
        k = by_kind.setdefault(st["kind"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        k["ms"] += st["ms"]
        k["flops"] += st["flops"]
        k["bytes"] += st["bytes"]
        k["launches"] += 1
    tc = {f: by_kind.get("gemm", {}).get(f, 0) + by_kind.get("conv", {}).get(f, 0) for f in ("ms", "flops", "bytes", "launches")}
    achieved = tc["flops"] / (tc["ms"] * 1e-3) / 1e12 if tc["ms"] > 0 else 0.0
    # DRAM traffic of the family: dram__bytes_read.sum + dram__bytes_write.sum summed over its launches in one step: a CONSTANT
    # read from the committed ncu capture of this same command (profiles/README.md), not measured in this run; null when the
    # capture is not there or is of another batch
    traffic, traffic_src = None, None
    for tname in ("r02_traffic.json", "r01_traffic.json"):
        tpath = ROOT / "profiles" / tname
        if tpath.exists() and B == BATCH:
            tj = json.loads(tpath.read_text())
            traffic = tj["family_dram_bytes_per_step"]
            traffic_src = f"profiles/{tname}: constant from the committed ncu capture (one step at batch 64), not measured in this run"
            break
    peak = pk["bf16_tflops_sustained"]
    roofline = {"bound": "tensor", "kernel": "tcgen05 GEMM / convolution kernels: tc_gemm_kernel, tc_mlp_kernel, tc_bneck_kernel, tc_bneck_halo_kernel, stem_kernel",
                "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per step over the family's launches (algorithmic: "
                                                    f"{tc['bytes']:.3e})",
                "traffic_source": traffic_src, "peak_source": f"{pk_kind} bf16_tflops_sustained",
                "launches_per_step": tc["launches"],
                "share_of_step": round(tc["ms"] / sum(med), 4),
                "whole_forward_tflops": round(value / world * GFLOP_PER_FRAME / 1e3, 1),
                "whole_forward_frac": round(value / world * GFLOP_PER_FRAME / 1e3 / peak, 4),
                "by_kind_ms": {k: round(v["ms"], 3) for k, v in by_kind.items()},
                "per_kernel": per_kernel_roofline(per_launch, pk)}
    if args.profile_out:
        Path(args.profile_out).write_text(json.dumps(per_launch, indent=0))

    launch_desc = "stream launches"
    if graph is not None:
        launch_desc = "cuda graph replay" + (" (NCCL all-reduce captured in the graph)" if graph_has_allreduce else
                                             (" + all-reduce after each replay" if world > 1 else ""))
    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world) | {"batch_per_gpu": B, "global_batch": B * world, "launch": launch_desc,
                                                **({"library_options": options} if options else {})},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "returns": "compacted per-frame detections (xywh f64, score, foot point f64, zone index, count) + per-frame zone "
                               "counts, copied to pinned host memory every step"},
            "e2e_objects": e2e_objects,
            "gpu_launches": int(launches), "detections_last_step": n_det,
            "roofline": roofline}
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_detect_run(steps=2, warmup=1, frames_per_step=4, h=H_IN, w=W_IN, n_zones=N_ZONES)
        line["cpu_baseline"] = {"value": round(r["value"], 4), "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                "legs": r["legs"], "sample": r["sample"]}
        try:
            line["parity"] = parity_block(out, host)
        except Exception as e:
            line["parity"] = {"error": f"{type(e).__name__}: {e}"}
    emit(line)
    ctx.close()


# --------------------------------------------------------------------------------------------------------------
# config 1: one 1280x720 frame, 4 zones, batch 1
# --------------------------------------------------------------------------------------------------------------
def run_config1(ctx: Ctx, det, steps: int) -> dict:
    torch = ctx.torch
    from office_person_detection_vit_b200.detection.synthetic import synthetic_frames
    from office_person_detection_vit_b200.pipeline import DetectCountPipeline
    from office_person_detection_vit_b200.scene import H_CONFIG, grid_zones
    from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer
    from office_person_detection_vit_b200.zone import ZoneClassifier

    pipe = DetectCountPipeline(det, HomographyTransformer(H_CONFIG, FloorMapConfig()), ZoneClassifier(grid_zones(4), allow_overlap=False))
    host = torch.from_numpy(synthetic_frames(1, 720, 1280, seed=0)).pin_memory()
    frame = host.to(ctx.dev)
    hist = torch.zeros(1, 5, dtype=torch.int32, device=ctx.dev)
    for _ in range(3):
        out = pipe.run_tensors(frame, hist=hist)
    torch.cuda.synchronize()
    try:
        graph = pipe.capture(frame, hist=hist, zero_hist=True)
    except Exception:
        graph = None
        torch.cuda.synchronize()

    # per-step events so that the L2 flush between steps is outside the measured time
    ms = []
    for _ in range(steps):
        ctx.flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graph is None:
            hist.zero_()
        out = graph() if graph is not None else pipe.run_tensors(frame, hist=hist)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    lat = statistics.median(ms)
    # end to end: pinned host frame -> device, step, counts + detections back
    w0 = time.perf_counter()
    for _ in range(steps):
        frame.copy_(host, non_blocking=True)
        o = graph() if graph is not None else pipe.run_tensors(frame, hist=hist)
        res = (o["det_xywh"].cpu(), o["det_score"].cpu(), o["n_keep"].cpu(), hist.cpu())
    e2e_s = (time.perf_counter() - w0) / steps
    pk, _ = peaks()
    return {"metric": "frames/sec DETR-R50 1280x720 (-> 750x1333) bf16 batch 1 + homography + 4 zones", "unit": UNIT,
            "value": round(1e3 / lat, 2), "ms_per_step": round(lat, 4), "steps": steps, "n_gpus": 1,
            "e2e": {"value": round(1.0 / e2e_s, 2), "unit": UNIT, "h2d_bytes_per_step": host.numel(),
                    "d2h_bytes_per_step": sum(t.numel() * t.element_size() for t in res)},
            "whole_forward_frac": round(1e3 / lat * GFLOP_PER_FRAME_720P / 1e3 / pk["bf16_tflops_sustained"], 4),
            "detections": int(out["n_keep"].sum().item()), "config": config1_workload() | {"launch": "cuda graph replay" if graph else "stream launches"}}


# --------------------------------------------------------------------------------------------------------------
# config 4: frame-sharded timelapse run, frames generated on the device, one all-reduce at the end
# --------------------------------------------------------------------------------------------------------------
def run_config4(ctx: Ctx, det, pipe, frames_per_gpu: int, graph: bool = True) -> dict:
    torch = ctx.torch
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection.vit_detector import synthetic_frames_device

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    B, Z = BATCH, N_ZONES
    T = frames_per_gpu * world
    first = rank * frames_per_gpu                    # this rank's contiguous shard: global frames [first, first + frames_per_gpu)
    n_full, tail = divmod(frames_per_gpu, B)
    buf = torch.empty(B, H_IN, W_IN, 3, dtype=torch.uint8, device=dev)
    hist = torch.zeros(T, Z + 1, dtype=torch.int32, device=dev)
    h64 = torch.zeros(B, Z + 1, dtype=torch.int32, device=dev)
    n_det = torch.zeros((), dtype=torch.int64, device=dev)

    synthetic_frames_device(buf, CONFIG4_SEED_BASE, first)
    pipe.run_tensors(buf, hist=h64, slot_base=0)                       # plan + warm-up (B = 64)
    if tail:
        pipe.run_tensors(buf[:tail], hist=torch.zeros(tail, Z + 1, dtype=torch.int32, device=dev), slot_base=0)   # plan of the tail batch
    g = None
    if graph:
        try:
            g = pipe.capture(buf, hist=h64, slot_base=0, zero_hist=True)
        except Exception as e:
            print(f"bench.py config 4: CUDA graph capture failed ({type(e).__name__}: {e}); plain launches", file=sys.stderr)
            g = None
            torch.cuda.synchronize()

    def run_all():
        hist.zero_()
        n_det.zero_()
        for i in range(n_full):
            g0 = first + i * B
            synthetic_frames_device(buf, CONFIG4_SEED_BASE, g0)          # no host I/O: the batch is generated where it is consumed
            if g is not None:
                o = g()                                                  # rows 0..63 of h64 = this batch's per-frame counts
                hist[g0:g0 + B].copy_(h64)
            else:
                o = pipe.run_tensors(buf, hist=hist, slot_base=g0)
            n_det.add_(o["n_keep"].sum())
        if tail:
            g0 = first + n_full * B
            synthetic_frames_device(buf[:tail], CONFIG4_SEED_BASE, g0)
            o = pipe.run_tensors(buf[:tail], hist=hist, slot_base=g0)
            n_det.add_(o["n_keep"].sum())
        pipe.all_reduce(hist)                                            # the run's ONE collective: [T, 17] int32

    launches0 = _lib.lib().opd_launch_count()
    sampler = ClockSampler(ctx.local)
    sampler.start()
    total_ms = ctx.timed(run_all, 1)
    clocks = sampler.finish()
    launches = _lib.lib().opd_launch_count() - launches0
    if g is not None:
        launches += n_full * (_graph_kernel_count(pipe, buf, h64))
    # checks (outside the timed region): rows of this rank's shard hold one count per kept detection; two batches re-run through
    # the plain (non-graph) API on regenerated frames give the same rows
    mine = hist[first:first + frames_per_gpu]
    ok_sum = int(mine.sum().item()) == int(n_det.item())
    ok_rows = True
    for i in sorted({0, max(n_full - 1, 0)}):
        g0 = first + i * B
        chk = torch.zeros(T, Z + 1, dtype=torch.int32, device=dev)
        synthetic_frames_device(buf, CONFIG4_SEED_BASE, g0)
        pipe.run_tensors(buf, hist=chk, slot_base=g0)
        ok_rows = ok_rows and bool(torch.equal(chk[g0:g0 + B], hist[g0:g0 + B]))
    total = int(hist.sum().item())                                       # after the all-reduce: every rank's detections
    return {"metric": "frames/sec DETR-R50 800x1333 bf16, frame-sharded timelapse run (detect + homography + 16 zones + per-frame counts)",
            "unit": UNIT, "value": round(world * frames_per_gpu / (total_ms / 1e3), 2), "total_ms": round(total_ms, 2),
            "ms_per_batch": round(total_ms / (n_full + (1 if tail else 0)), 4), "n_gpus": world, "scaling": "weak",
            "frames_total": world * frames_per_gpu, "hist_shape": [T, Z + 1], "collectives": 1, "gpu_launches": int(launches),
            "detections_total": total, "checks": {"hist_sum_equals_detections": ok_sum, "resampled_batches_equal": ok_rows},
            "clocks": clocks, "config": config4_workload(world, frames_per_gpu) |
            {"launch": ("cuda graph replay per batch + device copy of its 64 histogram rows" if g is not None else "stream launches")}}


def _graph_kernel_count(pipe, buf, h64) -> int:
    """Kernels of one captured step (counted once with plain launches)."""
    from office_person_detection_vit_b200 import _lib

    n0 = _lib.lib().opd_launch_count()
    pipe.run_tensors(buf, hist=h64, slot_base=0)
    return int(_lib.lib().opd_launch_count() - n0)


# --------------------------------------------------------------------------------------------------------------
# config 5: homography + point-in-polygon + zone count over 10^8 points (the HBM-roofline path)
# --------------------------------------------------------------------------------------------------------------
def run_config5(ctx: Ctx, n_points: int, steps: int, with_e2e: bool, with_cpu: bool = False) -> dict:
    torch = ctx.torch
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.scene import H_CONFIG, grid_zones
    from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer
    from office_person_detection_vit_b200.zone import ZoneClassifier

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    n = n_points // world
    Z = CONFIG5_ZONES
    gen = torch.Generator(device=dev).manual_seed(3 + rank)
    pts = torch.empty((n, 2), dtype=torch.float32, device=dev)
    pts[:, 0].uniform_(0, 1280, generator=gen)
    pts[:, 1].uniform_(0, 720, generator=gen)
    tr = HomographyTransformer(H_CONFIG, FloorMapConfig())
    zones = grid_zones(Z)
    zc = ZoneClassifier(zones, allow_overlap=False)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    hist = torch.zeros(1, Z + 1, dtype=torch.int32, device=dev)

    def step():
        hist.zero_()
        zc.count(pts, transformer=tr, out=hist, index_out=idx)
        if world > 1:
            ctx.dist.all_reduce(hist)

    for _ in range(3):
        step()
    launches0 = _lib.lib().opd_launch_count()
    sampler = ClockSampler(ctx.local)
    sampler.start()
    total_ms = ctx.timed(step, steps)
    launches = _lib.lib().opd_launch_count() - launches0
    value = n * world * steps / (total_ms / 1e3)

    def kernel_ms(fn, reps: int) -> float:
        """Median duration of the kernel alone: CUDA events on the launching stream around each launch."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
        torch.cuda.synchronize()
        for i in range(reps):
            ev[2 * i].record()
            fn()
            ev[2 * i + 1].record()
        torch.cuda.synchronize()
        return statistics.median(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(reps))

    pk, pk_kind = peaks()
    k_ms = kernel_ms(lambda: zc.count(pts, transformer=tr, out=hist, index_out=idx), max(steps, 10))
    clocks = sampler.finish()
    gbs = n * BYTES_PER_POINT / (k_ms * 1e-3) / 1e9
    variants = []
    for zv in (4, 16, 64):
        zcv = zc if zv == Z else ZoneClassifier(grid_zones(zv), allow_overlap=False)
        hv = torch.zeros(1, zv + 1, dtype=torch.int32, device=dev)
        for mode, fn, bpp in (("idx+count", lambda: zcv.count(pts, transformer=tr, out=hv, index_out=idx), 12),
                              ("count", lambda: zcv.count(pts, transformer=tr, out=hv), 8)):
            fn()
            m = kernel_ms(fn, 10)
            variants.append({"zones": zv, "mode": mode, "ms": round(m, 4), "gbs": round(n * bpp / (m * 1e-3) / 1e9, 1),
                             "frac_hbm": round(n * bpp / (m * 1e-3) / 1e9 / pk["hbm_gbs"], 4)})

    # parity (the oracle as the checker): the first 10^6 points of rank 0 against the C oracle; counts = histogram of the indices
    parity = None
    if rank == 0:
        import numpy as np

        from oracle import floor_oracle as fo

        hist.zero_()
        zc.count(pts, transformer=tr, out=hist, index_out=idx)
        torch.cuda.synchronize()
        m = min(n, 1_000_000)
        host_pts = pts[:m].cpu().numpy()
        exp_idx, _ = fo.project_classify_count(H_CONFIG, host_pts, zones)
        got = idx[:m].cpu().numpy()
        px, _, _ = fo.transform(H_CONFIG, host_pts.astype(np.float64), is_bbox=False)
        far = fo.min_edge_distance(px, zones) >= 1e-4
        bins = torch.bincount(torch.where(idx < 0, Z, idx).long(), minlength=Z + 1).to(torch.int32)
        parity = {"points_checked": m, "mismatches_outside_1e-4px_band": int((got[far] != exp_idx[far]).sum()),
                  "mismatches_total": int((got != exp_idx).sum()), "hist_equals_bincount_of_indices": bool(torch.equal(bins, hist[0])),
                  "hist_sum_equals_points": int(hist.sum().item()) == n}
    res = {"metric": METRIC5, "unit": UNIT5, "value": round(value, 1), "ms_per_step": round(total_ms / steps, 4), "steps": steps,
           "n_gpus": world, "scaling": "strong", "dtype": "f32 filter + f64 exact path", "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "floor_fast_kernel (projection + classification + count, idx + count mode)",
                        "achieved": round(gbs, 1), "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(gbs / pk["hbm_gbs"], 4),
                        "kernel_ms": round(k_ms, 4), "algorithmic_bytes_per_launch": n * BYTES_PER_POINT, "traffic": None,
                        "peak_source": f"{pk_kind} hbm_gbs", "variants": variants},
           "clocks": clocks, "parity": parity, "config": config5_workload(world, n_points)}
    if with_e2e:
        # end to end with HOST buffers: pinned points -> device, kernel, zone indices + counts back to pinned host memory
        h_pts = pts.cpu().pin_memory()
        h_idx = torch.empty(n, dtype=torch.int32).pin_memory()
        h_hist = torch.empty(1, Z + 1, dtype=torch.int32).pin_memory()
        d_pts = torch.empty_like(pts)

        def e2e_step():
            d_pts.copy_(h_pts, non_blocking=True)
            hist.zero_()
            zc.count(d_pts, transformer=tr, out=hist, index_out=idx)
            if world > 1:
                ctx.dist.all_reduce(hist)
            h_idx.copy_(idx, non_blocking=True)
            h_hist.copy_(hist, non_blocking=True)

        e2e_step()
        k = max(3, min(steps, 5))
        ctx.barrier()
        e2e_ms = ctx.timed(e2e_step, k)
        res["e2e"] = {"value": round(n * world * k / (e2e_ms / 1e3), 1), "unit": UNIT5, "h2d_bytes_per_step": n * 8,
                      "d2h_bytes_per_step": n * 4 + (Z + 1) * 4, "steps": k}
    if with_cpu and rank == 0:
        res["cpu_baseline"] = cpu_floor_run(1_000_000, Z)
    return res


def main_single_config(args) -> None:
    ctx = Ctx()
    if args.config == 5:
        res = run_config5(ctx, args.points, args.steps, with_e2e=True, with_cpu=not args.no_cpu_baseline)
    else:
        det, pipe, options = build_pipeline(ctx, N_ZONES, BATCH)
        if args.config == 4:
            res = run_config4(ctx, det, pipe, args.config4_frames, graph=not args.no_graph)
        else:
            res = run_config1(ctx, det, steps=args.steps)
            if ctx.rank == 0 and not args.no_cpu_baseline:
                r = cpu_detect_run(steps=3, warmup=1, frames_per_step=1, h=720, w=1280, n_zones=4)
                r.pop("first", None)
                res["cpu_baseline"] = {k: r[k] for k in ("value", "cores", "kind", "legs", "sample")} | {"unit": UNIT}
    if ctx.rank == 0:
        res.setdefault("steps", args.steps)
        res |= {"warmup": 3, "higher_is_better": True, "vs_baseline": None, "data": "synthetic"}
        res.setdefault("dtype", "bf16")
        emit(res)
    ctx.close()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line, on the process's original stdout."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


def main() -> None:
    # Libraries write banners to file descriptor 1 (NCCL prints its version there when NCCL_DEBUG is set on the box): keep the real
    # stdout for the one JSON line and point fd 1 at stderr for everything else.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5], help="BASELINE.json config (default 3: the headline)")
    ap.add_argument("--batch", type=int, default=BATCH, help="frames per GPU per step (the metric is quoted at 64)")
    ap.add_argument("--points", type=int, default=CONFIG5_POINTS, help="config 5: total points over all GPUs")
    ap.add_argument("--config4-frames", type=int, default=CONFIG4_FRAMES_PER_GPU, help="config 4: frames per GPU (12 500 = 100 000 / 8)")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip the extra.config1/4/5 sub-benchmarks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time plain stream launches instead of CUDA-graph replays")
    ap.add_argument("--no-graph-allreduce", action="store_true", help="keep the NCCL all-reduce outside the captured graph")
    ap.add_argument("--profile-out", default="", help="write the per-launch timing table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    elif args.config in (2, 3):
        main_headline(args)
    else:
        main_single_config(args)


if __name__ == "__main__":
    main()
