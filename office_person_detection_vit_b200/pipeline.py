"""Phase 2 -> 3 (+ count) on device tensors: detect -> foot point -> homography -> zone -> per-timestamp histogram.

The reference runs these as three Python loops (src/pipeline/phases/detection.py:56-133, transform.py:257-330,
aggregation.py:26-91) over lists of Detection objects.  Here the whole chain stays on the GPU: the detector's compacted
foot points feed the fused projection + classification + histogram kernel directly, one timestamp slot per frame, and the
only host traffic is whatever the caller reads back.  Multi-GPU: every rank owns a contiguous block of frames / slots
and `all_reduce(hist)` (NCCL) merges the per-timestamp histograms."""

from __future__ import annotations

from . import _lib


class DetectCountPipeline:
    def __init__(self, detector, transformer, zone_classifier):
        self.detector = detector
        self.transformer = transformer
        self.zones = zone_classifier

    def run_tensors(self, frames, hist=None, slot_base: int = 0, threshold: float | None = None, bgr: bool = True) -> dict:
        """frames [B,H,W,3] uint8 CUDA -> dict of device tensors: everything `ViTDetector.detect_tensors` returns plus
        `zone_idx` [B,100] (-1 = unclassified, and -1 for the unused rows >= n_keep[b]: their slot is -1 and the floor kernel
        neither classifies nor counts them) and `hist` [T, Z+1] int32, accumulated at rows
        slot_base .. slot_base+B-1 (allocated as [B, Z+1] when not given).  No host synchronisation."""
        torch = _lib.require_cuda()
        B = frames.shape[0]
        out = self.detector.detect_tensors(frames, bgr=bgr, threshold=threshold, slot_base=slot_base)
        if hist is None:
            hist = torch.zeros(slot_base + B, self.zones.get_zone_count() + 1, dtype=torch.int32, device=frames.device)
        pts = out["det_foot"].view(-1, 2)
        if hasattr(self.transformer, "floor_params"):
            # homography: projection, classification and counting are one fused kernel (csrc/floor.cu)
            hist, idx = self.zones.count(pts, slot=out["det_slot"].view(-1), num_slots=hist.shape[0],
                                         transformer=self.transformer, out=hist, return_index=True)
        else:
            # piecewise-affine (the reference's default transform.method): its own kernel (csrc/pwa.cu), then classification +
            # counting on the floor points without a projection
            out["floor_px"] = self.transformer.transform_points(pts)
            hist, idx = self.zones.count(out["floor_px"], slot=out["det_slot"].view(-1), num_slots=hist.shape[0], transformer=None,
                                         out=hist, return_index=True)
        out["hist"] = hist
        out["zone_idx"] = idx.view(B, -1)
        return out

    def capture(self, frames, hist=None, slot_base: int = 0, threshold: float | None = None, bgr: bool = True, zero_hist: bool = False,
                all_reduce: bool = False):
        """Capture one `run_tensors(frames, ...)` step into a CUDA graph and return a callable that replays it and returns the
        (static) output tensors: for steady-state loops that refill the SAME `frames` buffer (and `hist`) between steps.  The
        143 launches of a step are enqueued by one graph launch instead of 143 API calls.  `zero_hist` clears `hist` inside the
        graph before the step; `all_reduce` captures the NCCL all-reduce of `hist` as the graph's last node (multi-GPU: the step's
        only collective then costs no launch of its own).  Everything the step reads (frames, weights, tables) must stay where it is
        while the graph lives."""
        torch = _lib.require_cuda()
        if hist is None:
            hist = torch.zeros(slot_base + frames.shape[0], self.zones.get_zone_count() + 1, dtype=torch.int32, device=frames.device)
        self.run_tensors(frames, hist=hist, slot_base=slot_base, threshold=threshold, bgr=bgr)   # plans, kernel attributes, allocations
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            if zero_hist:
                hist.zero_()
            out = self.run_tensors(frames, hist=hist, slot_base=slot_base, threshold=threshold, bgr=bgr)
            if all_reduce:
                self.all_reduce(hist)

        def replay():
            graph.replay()
            return out

        replay.graph = graph
        return replay

    def run_stream(self, batches, hist=None, slot_base: int = 0, threshold: float | None = None, bgr: bool = True):
        """Frame ingest (the reference hands Phase 2 a Python list of cv2 frames one by one,
        src/pipeline/orchestrator.py:155-202, src/pipeline/phases/detection.py:91-94): `batches` yields host batches - uint8
        [B,H,W,3] NumPy arrays or torch tensors of one size - and this generator yields `run_tensors` results, with batch i+1
        staged into pinned memory and copied to the device on a second stream while batch i runs (two pinned and two device
        buffers).  Batch i's timestamp rows follow batch i - 1's (a running offset: batches may differ in size, e.g. a short
        last one); `hist` (if given) must hold every batch's rows."""
        torch = _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream()
        pinned, device_bufs = [None, None], [None, None]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        for c in consumed:
            c.record(main)

        def stage(i, batch):
            t = torch.as_tensor(batch)
            if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
                raise ValueError(f"batch {i}: expected uint8 [B,H,W,3], got {t.dtype} {tuple(t.shape)}")
            k = i % 2
            if device_bufs[k] is None or device_bufs[k].shape != t.shape:
                device_bufs[k] = torch.empty(t.shape, dtype=torch.uint8, device=dev)
            consumed[k].synchronize()          # batch i - 2 has been consumed: device_bufs[k] (and pinned[k]) may be overwritten
            if t.is_pinned():
                src = t                        # the producer already writes pinned memory (bench.py): no staging copy
            else:
                if pinned[k] is None or pinned[k].shape != t.shape:
                    pinned[k] = torch.empty(t.shape, dtype=torch.uint8).pin_memory()
                pinned[k].copy_(t)
                src = pinned[k]
            with torch.cuda.stream(copy_stream):
                device_bufs[k].copy_(src, non_blocking=True)
                ready[k].record(copy_stream)

        it = iter(batches)
        nxt = next(it, None)
        i = 0
        slot = slot_base
        if nxt is not None:
            stage(0, nxt)
        while nxt is not None:
            cur_k = i % 2
            nxt = next(it, None)
            if nxt is not None:
                stage(i + 1, nxt)
            main.wait_event(ready[cur_k])
            B = device_bufs[cur_k].shape[0]
            out = self.run_tensors(device_bufs[cur_k], hist=hist, slot_base=slot, threshold=threshold, bgr=bgr)
            consumed[cur_k].record(main)
            yield out
            i += 1
            slot += B

    def all_reduce(self, hist):
        """Sum the per-timestamp histograms over all ranks (the path's only collective)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        return hist
