#!/usr/bin/env python
"""Benchmark of the Phase 2 -> 3 hot path (BASELINE.json): frames/sec of DETR-ResNet-50 at 800x1333, bf16, batch 64
per GPU, followed by foot-point homography + 16-zone classification + per-timestamp counting.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's CPU path on the host cores)

One "step" = one pass of the hot path over one batch of 64 synthetic frames per rank (weak scaling: frames shard
across ranks, the only collective is the NCCL all-reduce of the [T, Z+1] zone-count histogram).
Prints ONE JSON line (rank 0).  See DESIGN.md "measurement" for what each field means.
"""

from __future__ import annotations

import argparse
import json
import os
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
GFLOP_PER_FRAME = 203.18          # algorithmic, 2 flops per MAC, convs + linears + attention matmuls
H_IN, W_IN, BATCH, N_ZONES = 800, 1333, 64, 16


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

    def summary(self) -> dict:
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 6 and s[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU implementation of the path on the host cores
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, frames_per_step: int = 1) -> dict:
    """transformers' DetrForObjectDetection + DetrImageProcessor in float32 on all host threads (the arithmetic the
    reference's removed ViTDetector drove; random-init weights, synthetic 800x1333 frames), then the reference's
    transform_batch + classify + get_zone_counts as restated by oracle/floor_oracle (the reference's own sources are
    not on the GPU box).  Returns frames/s over `steps` steps of `frames_per_step` frames."""
    import numpy as np
    import torch

    from office_person_detection_vit_b200.detection.synthetic import random_init_state_dict, synthetic_frames
    from office_person_detection_vit_b200.scene import H_CONFIG, grid_zones
    from oracle import detr_oracle as do
    from oracle import floor_oracle as fo

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = random_init_state_dict(0)
    model = do.hf_model(w)
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    from transformers import DetrImageProcessor

    proc = DetrImageProcessor()
    zones = grid_zones(N_ZONES)
    frames = synthetic_frames(frames_per_step, H_IN, W_IN, seed=1)

    def one_step():
        rgb = [np.ascontiguousarray(f[:, :, ::-1]) for f in frames]
        with torch.no_grad():
            inp = proc(images=rgb, return_tensors="pt")
            out = model(**inp)
            res = proc.post_process_object_detection(out, threshold=0.5,
                                                     target_sizes=torch.tensor([[H_IN, W_IN]] * len(rgb)))
        n = 0
        for r in res:
            keep = r["labels"] == do.PERSON_LABEL
            b = r["boxes"][keep].double().numpy()
            if len(b):
                xywh = np.stack([b[:, 0], b[:, 1], b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], axis=1)
                px, _, _ = fo.transform(H_CONFIG, xywh, is_bbox=True)
                idx, _ = fo.classify(px, zones)
                fo.count(zone_idx=idx, Z=N_ZONES)
                n += len(b)
        return n

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    return {"value": steps * frames_per_step / dt, "ms_per_step": dt / steps * 1e3, "cores": cores,
            "sample": f"{steps} steps x {frames_per_step} synthetic 800x1333 frame(s): transformers DetrImageProcessor + "
                      f"DetrForObjectDetection fp32 eager on {cores} threads + oracle port of transform_batch/classify/"
                      f"get_zone_counts (16 zones)"}


def main_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(max(1, args.steps), max(1, min(args.warmup, 1)), frames_per_step=1)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus: int) -> dict:
    return {"workload": "BASELINE configs[1] DETR-ResNet-50 (random-init) bf16 batch=64 at 800x1333 detection per GPU, "
                        "followed by the configs[2] tail (foot-point homography + 16-polygon zone classification + "
                        "per-frame counts [64,17]); frames sharded across GPUs, one NCCL all-reduce of the histogram per step",
            "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "frame": [H_IN, W_IN, 3], "zones": N_ZONES,
            "cache": "inputs larger than L2 (205 MB of frames, GBs of activations per step)",
            "parallelism": f"frames dp{n_gpus}"}


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def main_gpu(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ViTDetector
    from office_person_detection_vit_b200.detection.synthetic import random_init_state_dict, synthetic_frames
    from office_person_detection_vit_b200.pipeline import DetectCountPipeline
    from office_person_detection_vit_b200.scene import H_CONFIG, grid_zones
    from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer
    from office_person_detection_vit_b200.zone import ZoneClassifier

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set on the box: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # development A/B switch: OPD_OPTIONS="name=value,..." -> opd_set_option before the plans are built; recorded in `config`
    options = dict(kv.split("=") for kv in os.environ.get("OPD_OPTIONS", "").split(",") if kv)
    for k, v in options.items():
        _lib.check(_lib.lib().opd_set_option(k.encode(), int(v)), f"opd_set_option({k})")
    det = ViTDetector(confidence_threshold=0.5, state_dict=random_init_state_dict(0), device=f"cuda:{local}",
                      batch_size=BATCH)
    det.load_model()
    pipe = DetectCountPipeline(det, HomographyTransformer(H_CONFIG, FloorMapConfig()),
                               ZoneClassifier(grid_zones(N_ZONES), allow_overlap=False))
    B = args.batch
    host = torch.from_numpy(synthetic_frames(B, H_IN, W_IN, seed=1 + rank)).pin_memory()
    frames = host.to(dev)
    hist = torch.zeros(B * world, N_ZONES + 1, dtype=torch.int32, device=dev)

    def step(fr):
        hist.zero_()
        out = pipe.run_tensors(fr, hist=hist, slot_base=rank * B)
        pipe.all_reduce(hist)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(frames)
    barrier()

    # ---- device-resident throughput ----
    # The step's launches (143: all on one stream, tensor maps pre-encoded, no host synchronisation) are captured ONCE into a
    # CUDA graph and replayed per step; the all-reduce of the histogram stays outside the graph.  --no-graph times plain launches.
    launches0 = _lib.lib().opd_launch_count()
    out = step(frames)
    launches_per_step = _lib.lib().opd_launch_count() - launches0
    graph = None
    if not args.no_graph:
        try:
            graph = pipe.capture(frames, hist=hist, slot_base=rank * B, zero_hist=True)   # the library's own capture API
        except Exception as e:   # capture is an optimisation, never a requirement
            print(f"bench.py: CUDA graph capture failed ({type(e).__name__}: {e}); timing plain launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def timed_step():
        if graph is None:
            return step(frames)
        o = graph()
        pipe.all_reduce(hist)
        return o

    for _ in range(2):
        timed_step()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = timed_step()
    e1.record()
    barrier()
    launches = launches_per_step * args.steps
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms / 1e3)
    n_det = int(out["n_keep"].sum().item())

    # ---- end to end through the public API with host buffers (pinned), double-buffered H2D on a copy stream ----
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [torch.empty_like(frames), torch.empty_like(frames)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    results = []

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            dev_bufs[i % 2].copy_(host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    e2e_graphs = [None, None]
    if graph is not None:
        try:
            e2e_graphs = [pipe.capture(dev_bufs[k], hist=hist, slot_base=rank * B, zero_hist=True) for k in range(2)]
        except Exception as e:
            print(f"bench.py: CUDA graph capture (e2e) failed ({type(e).__name__}: {e}); plain launches", file=sys.stderr)
            e2e_graphs = [None, None]
            torch.cuda.synchronize()

    def e2e_step(k):
        if e2e_graphs[k] is None:
            return step(dev_bufs[k])
        o = e2e_graphs[k]()
        pipe.all_reduce(hist)
        return o

    def e2e_run(n):
        nonlocal results
        results = []
        for c in consumed:
            c.record()
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            o = e2e_step(i % 2)
            consumed[i % 2].record()
            # the step's result: per-frame zone counts + detection count (reference: FrameResult.zone_counts)
            results.append((hist[rank * B:(rank + 1) * B].to("cpu", non_blocking=True), o["n_keep"].to("cpu", non_blocking=True)))
        torch.cuda.synchronize()

    e2e_run(2)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    e2e_run(args.steps)
    t1.record()
    barrier()
    wall = time.perf_counter() - w0
    e2e_ms = torch.tensor([max(t0.elapsed_time(t1), wall * 1e3)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(e2e_ms.item()) / 1e3)
    h2d = host.numel()
    d2h = B * (N_ZONES + 1) * 4 + B * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tc_gemm_kernel: every convolution and linear layer), CUDA events between launches ----
    pk, pk_kind = peaks()
    pipe.run_tensors(frames, hist=hist, slot_base=rank * B)   # local only: the other ranks have left, no collective here
    torch.cuda.synchronize()
    profs = [det.model.profile() for _ in range(3)]
    n_steps = len(profs[0])
    med = [statistics.median(p[i]["ms"] for p in profs) for i in range(n_steps)]
    by_kind: dict[str, dict] = {}
    for i, st in enumerate(profs[0]):
        k = by_kind.setdefault(st["kind"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        k["ms"] += med[i]
        k["flops"] += st["flops"]
        k["bytes"] += st["bytes"]
        k["launches"] += 1
    tc_ms = by_kind.get("gemm", {"ms": 0})["ms"] + by_kind.get("conv", {"ms": 0})["ms"]
    tc_flops = by_kind.get("gemm", {"flops": 0})["flops"] + by_kind.get("conv", {"flops": 0})["flops"]
    tc_bytes = by_kind.get("gemm", {"bytes": 0})["bytes"] + by_kind.get("conv", {"bytes": 0})["bytes"]
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    # DRAM traffic of the family: dram__bytes_read.sum + dram__bytes_write.sum summed over its launches in one step, from the
    # committed ncu capture of this same command (profiles/README.md); null when the capture is not there or is of another batch
    traffic, traffic_src = None, None
    tpath = Path(__file__).resolve().parent / "profiles" / "r01_traffic.json"
    if tpath.exists() and B == BATCH:
        tj = json.loads(tpath.read_text())
        traffic, traffic_src = tj["family_dram_bytes_per_step"], "profiles/r01_traffic.json (ncu, one step at batch 64)"
    peak = pk["bf16_tflops_sustained"]
    roofline = {"bound": "tensor", "kernel": "tcgen05 GEMM / convolution kernels: tc_gemm_kernel, tc_bneck_kernel, tc_bneck_halo_kernel, stem_kernel",
                "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per step over the family's launches (algorithmic: "
                                                    f"{tc_bytes:.3e})",
                "traffic_source": traffic_src, "peak_source": f"{pk_kind} bf16_tflops_sustained",
                "launches_per_step": by_kind.get("gemm", {"launches": 0})["launches"] + by_kind.get("conv", {"launches": 0})["launches"],
                "share_of_step": round(tc_ms / sum(med), 4),
                "whole_forward_tflops": round(value / world * GFLOP_PER_FRAME / 1e3, 1),
                "whole_forward_frac": round(value / world * GFLOP_PER_FRAME / 1e3 / peak, 4),
                "by_kind_ms": {k: round(v["ms"], 3) for k, v in by_kind.items()}}
    if args.profile_out:
        Path(args.profile_out).write_text(json.dumps(
            [{**st, "ms": med[i]} for i, st in enumerate(profs[0])], indent=0))

    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world) | {"batch_per_gpu": B, "global_batch": B * world,
                                                "launch": "cuda graph replay" if graph is not None else "stream launches",
                                                **({"library_options": options} if options else {})},
            "clocks": sampler.summary(),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "detections_last_step": n_det,
            "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=3, warmup=1, frames_per_step=1)
        line["cpu_baseline"] = {"value": round(r["value"], 4), "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": r["sample"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="frames per GPU per step (the metric is quoted at 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time plain stream launches instead of CUDA-graph replays")
    ap.add_argument("--profile-out", default="", help="write the per-launch timing table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
