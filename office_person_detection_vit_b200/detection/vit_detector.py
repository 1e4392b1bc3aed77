"""DETR-ResNet-50 person detector on the GPU, with the call surface of the reference's detector.

Surface kept: the removed `src/detection/vit_detector.py` ViTDetector
(`__init__(model_name, confidence_threshold, device)`, `load_model`, `detect`, `detect_batch`, `_get_foot_position`,
attributes `model`, `device`, `confidence_threshold`; method table in the reference's coverage.json) and its surviving
drop-in twin `src/detection/yolov8_detector.py:26-254` (same signatures, `RuntimeError("Model not loaded. Call
load_model() first.")` before load, `Detection(bbox=(x,y,w,h), confidence, class_id, class_name="person",
camera_coords=foot)` records).

All arithmetic runs in libopd_b200.so (csrc/detr_engine.cu): preprocessing, ResNet-50, transformer, heads and
post-processing are hand-written sm_100a kernels; torch tensors are device buffers only.  There is no CPU path.

Tensor entries (north star): `forward_raw(frames_u8)` -> (logits [B,100,92], boxes [B,100,4]) and
`detect_tensors(frames_u8)` -> compacted per-frame person detections as device tensors (no Python objects).
"""

from __future__ import annotations

import ctypes as C
import logging
from pathlib import Path
from typing import Sequence

import numpy as np

from .. import _lib
from ..models import Detection

logger = logging.getLogger(__name__)

_P = C.c_void_p


class _TensorF32(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


_lib.register("opd_detr_create", C.c_int, [C.POINTER(_TensorF32), C.c_int32, C.c_int32, C.POINTER(_P)])
_lib.register("opd_detr_destroy", None, [_P])
_lib.register("opd_detr_set_debug", C.c_int, [_P, C.c_int32])
_lib.register("opd_detr_set_resize", C.c_int, [_P, C.c_int32])
_lib.register("opd_detr_set_fusion", C.c_int, [_P, C.c_int32])
_lib.register("opd_detr_input_shape", C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                               C.POINTER(C.c_int32), C.POINTER(C.c_int32)])
_lib.register("opd_detr_workspace_bytes", C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)])
_lib.register("opd_detr_forward", C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_size_t, _P, _P,
                                           _P])
_lib.register("opd_detr_profile", C.c_int, [_P, _P, C.c_int32, C.POINTER(C.c_int32), _P, _P, _P, _P, _P, C.c_int32])
_lib.register("opd_detr_tap", C.c_int, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int32)])
_lib.register("opd_detr_tap_copy", C.c_int, [_P, C.c_char_p, _P, C.c_size_t, _P])
_lib.register("opd_roi_features_bf16", C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32, C.c_int32,
                                                C.c_int32, _P, _P])
_lib.register("opd_detr_postprocess", C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                               C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P])

class _FrameGroup(C.Structure):
    _fields_ = [("frames_dev", C.c_void_p), ("n", C.c_int32), ("H0", C.c_int32), ("W0", C.c_int32), ("frames_are_bgr", C.c_int32)]


_lib.register("opd_detr_workspace_bytes_mixed", C.c_int, [_P, C.POINTER(_FrameGroup), C.c_int32, C.POINTER(C.c_size_t)])
_lib.register("opd_detr_forward_mixed", C.c_int, [_P, C.POINTER(_FrameGroup), C.c_int32, _P, C.c_size_t, _P, _P, _P])
_lib.register("opd_synthetic_frames_u8", C.c_int, [C.c_uint64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, _P])

N_QUERIES = 100
N_LOGITS = 92
PERSON_LABEL = 1   # COCO id of "person" in facebook/detr-resnet-50 (reference tests use class_id=1)

# transformers 4.x state-dict names -> the 5.x names the library expects (both are accepted by load_model)
_KEY_RENAMES = (
    ("model.backbone.conv_encoder.model.", "model.backbone.model."),
    (".fc1.", ".mlp.fc1."),
    (".fc2.", ".mlp.fc2."),
    (".self_attn.out_proj.", ".self_attn.o_proj."),
    (".encoder_attn.out_proj.", ".encoder_attn.o_proj."),
)


def input_shape(h0: int, w0: int) -> tuple[int, int, int, int]:
    """(H_in, W_in, h_feat, w_feat): model input size (800/1333 rule) and stage-4 feature map of an h0 x w0 frame."""
    out = [C.c_int32() for _ in range(4)]
    _lib.check(_lib.lib().opd_detr_input_shape(h0, w0, *[C.byref(o) for o in out]), "opd_detr_input_shape")
    return tuple(o.value for o in out)


def input_shape_noresize(h0: int, w0: int) -> tuple[int, int]:
    """Stage-4 feature map of an h0 x w0 model input (every stride-2 layer: out = (in - 1) // 2 + 1, five times)."""
    for _ in range(5):
        h0, w0 = (h0 - 1) // 2 + 1, (w0 - 1) // 2 + 1
    return h0, w0


class DetrEngine:
    """Handle of the device-resident model + a cached workspace per (B, H0, W0)."""

    def __init__(self, state_dict: dict, device_index: int | None = None):
        torch = _lib.require_cuda()
        self.device_index = torch.cuda.current_device() if device_index is None else int(device_index)
        names, arrays = [], []
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked"):
                continue
            if ".mlp." not in k:
                for old, new in _KEY_RENAMES:
                    if old in k:
                        k = k.replace(old, new)
            arr = v.detach().to("cpu", torch.float32).contiguous().numpy() if hasattr(v, "detach") else \
                np.ascontiguousarray(v, dtype=np.float32)
            names.append(k.encode())
            arrays.append(arr)
        tens = (_TensorF32 * len(names))()
        for i, (n, a) in enumerate(zip(names, arrays)):
            tens[i].name, tens[i].data, tens[i].numel = n, a.ctypes.data, a.size
        handle = _P()
        _lib.check(_lib.lib().opd_detr_create(tens, len(names), self.device_index, C.byref(handle)), "opd_detr_create")
        self._h = handle.value
        self._ws: dict[tuple[int, int, int], object] = {}
        self._torch = torch
        self.do_resize = True

    def set_debug(self, on: bool) -> None:
        _lib.check(_lib.lib().opd_detr_set_debug(self._h, int(on)), "opd_detr_set_debug")
        self._ws.clear()

    def set_resize(self, on: bool) -> None:
        """False: frames are fed at their own size, like DetrImageProcessor(do_resize=False)."""
        _lib.check(_lib.lib().opd_detr_set_resize(self._h, int(on)), "opd_detr_set_resize")
        self.do_resize = bool(on)
        self._ws.clear()

    def set_fusion(self, on: bool) -> None:
        """False: stages 1-2 run their 3x3 and 1x1-expansion convolutions as separate kernels (A/B comparison)."""
        _lib.check(_lib.lib().opd_detr_set_fusion(self._h, int(on)), "opd_detr_set_fusion")

    def workspace(self, B: int, H0: int, W0: int):
        key = (B, H0, W0)
        ws = self._ws.get(key)
        if ws is None:
            n = C.c_size_t()
            _lib.check(_lib.lib().opd_detr_workspace_bytes(self._h, B, H0, W0, C.byref(n)), "opd_detr_workspace_bytes")
            ws = self._torch.empty(n.value + 1024, dtype=self._torch.uint8, device=f"cuda:{self.device_index}")
            self._ws[key] = ws
        return ws

    def forward(self, frames, bgr: bool = True, logits=None, boxes=None):
        """frames [B,H0,W0,3] uint8 CUDA tensor -> (logits [B,100,92] f32, boxes [B,100,4] f32 cxcywh in [0,1])."""
        torch = self._torch
        if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3
                and frames.is_contiguous()):
            raise ValueError("frames must be a contiguous [B,H,W,3] uint8 CUDA tensor")
        B, H0, W0, _ = frames.shape
        ws = self.workspace(B, H0, W0)
        base = (ws.data_ptr() + 1023) & ~1023
        if logits is None:
            logits = torch.empty(B, N_QUERIES, N_LOGITS, dtype=torch.float32, device=frames.device)
        if boxes is None:
            boxes = torch.empty(B, N_QUERIES, 4, dtype=torch.float32, device=frames.device)
        with torch.cuda.device(self.device_index):   # the stream of the ENGINE's device, whatever the caller's current device is
            rc = _lib.lib().opd_detr_forward(self._h, frames.data_ptr(), B, H0, W0, int(bgr), base,
                                             ws.numel() - (base - ws.data_ptr()), logits.data_ptr(), boxes.data_ptr(),
                                             _lib.stream_ptr())
        _lib.check(rc, "opd_detr_forward")
        return logits, boxes

    def forward_mixed(self, groups, bgr: bool = True):
        """A batch that mixes frame sizes, like DetrImageProcessor + DetrForObjectDetection on a list of images: `groups` is a list
        of contiguous [n, H, W, 3] uint8 CUDA tensors (one frame size per tensor).  Every frame is resized on its own, padded to the
        batch maximum with pixel_mask semantics (opd_detr_forward_mixed) -> (logits [B,100,92], boxes [B,100,4]) in group order."""
        torch = self._torch
        for g in groups:
            if not (g.is_cuda and g.dtype == torch.uint8 and g.dim() == 4 and g.shape[-1] == 3 and g.is_contiguous() and g.shape[0] > 0):
                raise ValueError("every group must be a non-empty contiguous [n,H,W,3] uint8 CUDA tensor")
        arr = (_FrameGroup * len(groups))()
        for i, g in enumerate(groups):
            arr[i].frames_dev, arr[i].n, arr[i].H0, arr[i].W0, arr[i].frames_are_bgr = g.data_ptr(), g.shape[0], g.shape[1], g.shape[2], int(bgr)
        key = ("mixed", tuple((g.shape[0], g.shape[1], g.shape[2]) for g in groups))
        ws = self._ws.get(key)
        if ws is None:
            n = C.c_size_t()
            _lib.check(_lib.lib().opd_detr_workspace_bytes_mixed(self._h, arr, len(groups), C.byref(n)), "opd_detr_workspace_bytes_mixed")
            ws = self._ws[key] = torch.empty(n.value + 1024, dtype=torch.uint8, device=f"cuda:{self.device_index}")
        base = (ws.data_ptr() + 1023) & ~1023
        B = sum(g.shape[0] for g in groups)
        dev = groups[0].device
        logits = torch.empty(B, N_QUERIES, N_LOGITS, dtype=torch.float32, device=dev)
        boxes = torch.empty(B, N_QUERIES, 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(self.device_index):
            rc = _lib.lib().opd_detr_forward_mixed(self._h, arr, len(groups), base, ws.numel() - (base - ws.data_ptr()),
                                                   logits.data_ptr(), boxes.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "opd_detr_forward_mixed")
        return logits, boxes

    STEP_KINDS = ("elementwise", "gemm", "conv", "attention", "heads")

    def profile(self) -> list[dict]:
        """One more forward with the last forward's arguments, a CUDA event between consecutive launches.
        -> [{name, kind, ms, flops, bytes}] per launch (algorithmic flops / bytes).  Synchronises."""
        n = C.c_int32()
        with self._torch.cuda.device(self.device_index):
            _lib.check(_lib.lib().opd_detr_profile(self._h, _lib.stream_ptr(), 0, C.byref(n), None, None, None, None, None, 0),
                       "opd_detr_profile")
            k = n.value
            kinds, flops, nbytes, ms = (C.c_int32 * k)(), (C.c_double * k)(), (C.c_double * k)(), (C.c_float * k)()
            names = C.create_string_buffer(k * 48)
            _lib.check(_lib.lib().opd_detr_profile(self._h, _lib.stream_ptr(), k, C.byref(n), kinds, flops, nbytes, ms, names, 48),
                       "opd_detr_profile")
        raw = names.raw
        return [{"name": raw[i * 48:(i + 1) * 48].split(b"\0", 1)[0].decode(), "kind": self.STEP_KINDS[kinds[i]],
                 "ms": float(ms[i]), "flops": float(flops[i]), "bytes": float(nbytes[i])} for i in range(k)]

    def tap(self, name: str):
        """Copy of a named internal activation of the last forward (needs set_debug(True) for the reused ones)."""
        torch = self._torch
        p, rows, cols, f32 = _P(), C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(_lib.lib().opd_detr_tap(self._h, name.encode(), C.byref(p), C.byref(rows), C.byref(cols), C.byref(f32)),
                   "opd_detr_tap")
        dt = {0: torch.bfloat16, 1: torch.float32, 2: torch.uint8}[f32.value]
        out = torch.empty(rows.value, cols.value, dtype=dt, device=f"cuda:{self.device_index}")
        with torch.cuda.device(self.device_index):
            _lib.check(_lib.lib().opd_detr_tap_copy(self._h, name.encode(), out.data_ptr(), out.numel() * out.element_size(),
                                                    _lib.stream_ptr()), "opd_detr_tap_copy")
        return out

    def roi_features(self, det_xywh, n_keep, h0: int, w0: int):
        """ROI mean-pool + L2 norm of the LAST forward's encoder output over the compacted boxes
        det_xywh [B,Q,4] f64 / n_keep [B] i32 (device) -> [B,Q,256] f32 (rows >= n_keep are zero).  No host sync."""
        torch = self._torch
        p, rows, cols, kind = _P(), C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(_lib.lib().opd_detr_tap(self._h, b"enc5", C.byref(p), C.byref(rows), C.byref(cols), C.byref(kind)), "opd_detr_tap")
        B, Q = det_xywh.shape[0], det_xywh.shape[1]
        _, _, fh, fw = input_shape(h0, w0) if self.do_resize else (0, 0, *input_shape_noresize(h0, w0))
        if rows.value != B * fh * fw:
            raise ValueError(f"roi_features: the last forward was not a batch of {B} frames of {h0}x{w0}")
        if not (det_xywh.is_cuda and det_xywh.dtype == torch.float64 and det_xywh.is_contiguous() and n_keep.dtype == torch.int32):
            raise ValueError("roi_features: det_xywh must be a contiguous float64 CUDA tensor, n_keep int32")
        out = torch.empty(B, Q, cols.value, dtype=torch.float32, device=det_xywh.device)
        with torch.cuda.device(self.device_index):
            rc = _lib.lib().opd_roi_features_bf16(p.value, B, fh, fw, cols.value, det_xywh.data_ptr(), n_keep.data_ptr(), Q, h0, w0,
                                                  out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "opd_roi_features_bf16")
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().opd_detr_destroy(self._h)
        except Exception:  # interpreter shutdown
            pass


def synthetic_frames_device(out, seed_base: int, global_frame0: int):
    """Fill `out` ([B,H,W,3] uint8 CUDA tensor) with synthetic frames generated on the device: frame b is global frame
    global_frame0 + b of the stream seeded by seed_base (opd_synthetic_frames_u8; bench.py config 4).  No host sync."""
    torch = _lib.require_cuda()
    if not (out.is_cuda and out.dtype == torch.uint8 and out.dim() == 4 and out.shape[-1] == 3 and out.is_contiguous()):
        raise ValueError("out must be a contiguous [B,H,W,3] uint8 CUDA tensor")
    B, H, W, _ = out.shape
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().opd_synthetic_frames_u8(int(seed_base), int(global_frame0), B, H, W, out.data_ptr(), _lib.stream_ptr()),
                   "opd_synthetic_frames_u8")
    return out


def postprocess_tensors(logits, boxes, h0: int, w0: int, threshold: float, person_label: int = PERSON_LABEL,
                        slot_base: int = 0) -> dict:
    """K8b on device tensors: per-query scores / labels / xyxy and the compacted person detections per frame."""
    torch = _lib.require_cuda()
    B, Q, Cn = logits.shape
    dev = logits.device
    out = {
        "scores": torch.empty(B, Q, dtype=torch.float32, device=dev),
        "labels": torch.empty(B, Q, dtype=torch.int32, device=dev),
        "xyxy": torch.empty(B, Q, 4, dtype=torch.float32, device=dev),
        "det_xywh": torch.zeros(B, Q, 4, dtype=torch.float64, device=dev),
        "det_score": torch.zeros(B, Q, dtype=torch.float32, device=dev),
        "det_foot": torch.zeros(B, Q, 2, dtype=torch.float64, device=dev),
        "det_query": torch.full((B, Q), -1, dtype=torch.int32, device=dev),
        "n_keep": torch.empty(B, dtype=torch.int32, device=dev),
        "det_slot": torch.empty(B, Q, dtype=torch.int32, device=dev),
    }
    with torch.cuda.device(dev):
        rc = _lib.lib().opd_detr_postprocess(
            logits.data_ptr(), boxes.data_ptr(), B, Q, Cn, h0, w0, float(threshold), person_label,
            out["scores"].data_ptr(), out["labels"].data_ptr(), out["xyxy"].data_ptr(), out["det_xywh"].data_ptr(),
            out["det_score"].data_ptr(), out["det_foot"].data_ptr(), out["det_query"].data_ptr(), out["n_keep"].data_ptr(),
            out["det_slot"].data_ptr(), int(slot_base), _lib.stream_ptr())
    _lib.check(rc, "opd_detr_postprocess")
    return out


class ViTDetector:
    """Person detector with the reference's ViTDetector / YOLOv8Detector surface, backed by libopd_b200.so."""

    def __init__(self, model_name: str = "facebook/detr-resnet-50", confidence_threshold: float = 0.5,
                 device: str | None = None, state_dict: dict | None = None, batch_size: int = 64, mixed_sizes: str = "pad"):
        self.model_name = model_name
        self.confidence_threshold = confidence_threshold
        self.device = self._setup_device(device)
        self.batch_size = int(batch_size)
        if mixed_sizes not in ("pad", "group"):
            raise ValueError("mixed_sizes must be 'pad' (one padded device batch, the reference's DetrImageProcessor behaviour) or "
                             "'group' (one device batch per frame size)")
        self.mixed_sizes = mixed_sizes
        self.model: DetrEngine | None = None
        self._state_dict = state_dict
        self._pinned: dict[tuple[int, int], object] = {}
        self._copy_pool = None
        logger.info(f"ViTDetector initialized with model: {model_name}")
        logger.info(f"Using device: {self.device}")
        logger.info(f"Confidence threshold: {confidence_threshold}")

    def _setup_device(self, device: str | None = None) -> str:
        """The reference falls back mps -> cuda -> cpu; this implementation is CUDA only and says so loudly."""
        if device is None:
            return "cuda"
        if not str(device).startswith("cuda"):
            raise ValueError(f"device={device!r}: office_person_detection_vit_b200 runs on CUDA (sm_100a) only")
        return str(device)

    def _device_index(self) -> int:
        torch = _lib.require_cuda()
        return int(self.device.split(":")[1]) if ":" in self.device else torch.cuda.current_device()

    def load_model(self) -> None:
        """Builds the device copy of the weights.  Weights come from `state_dict=` (transformers DETR key names) or,
        when `model_name` is a local file / directory, from `pytorch_model.bin` / `*.pt` / `*.safetensors` in it.
        There is no network in this environment, so a hub id without local weights raises RuntimeError."""
        try:
            sd = self._state_dict
            if sd is None:
                sd = self._load_local_state_dict(Path(self.model_name))
            self.model = DetrEngine(sd, self._device_index())
            logger.info(f"Model loaded: {self.model_name}")
        except Exception as e:
            logger.error(f"Failed to load model: {e}")
            raise RuntimeError(f"Failed to load DETR model: {e}") from e

    @staticmethod
    def _load_local_state_dict(path: Path) -> dict:
        import torch

        if path.is_dir():
            for name in ("model.safetensors", "pytorch_model.bin"):
                if (path / name).exists():
                    path = path / name
                    break
        if not path.is_file():
            raise FileNotFoundError(f"no local weights at {path} (pass state_dict= or a local checkpoint; no network access)")
        if path.suffix == ".safetensors":
            from safetensors.torch import load_file

            return load_file(str(path))
        sd = torch.load(str(path), map_location="cpu", weights_only=True)
        return sd.get("state_dict", sd) if isinstance(sd, dict) else sd

    # ---- tensor entries ------------------------------------------------------------------------------------
    def forward_raw(self, frames, bgr: bool = True):
        if self.model is None:
            raise RuntimeError("Model not loaded. Call load_model() first.")
        return self.model.forward(frames, bgr=bgr)

    def detect_tensors(self, frames, bgr: bool = True, threshold: float | None = None, slot_base: int = 0) -> dict:
        """frames [B,H,W,3] uint8 CUDA tensor -> dict of device tensors (see postprocess_tensors); no host sync."""
        logits, boxes = self.forward_raw(frames, bgr=bgr)
        _, H0, W0, _ = frames.shape
        thr = self.confidence_threshold if threshold is None else threshold
        out = postprocess_tensors(logits, boxes, H0, W0, thr, slot_base=slot_base)
        out["logits"], out["boxes"] = logits, boxes
        return out

    # ---- reference surface ---------------------------------------------------------------------------------
    def detect(self, frame: np.ndarray) -> list[Detection]:
        """frame: BGR uint8 ndarray [H,W,3] -> list[Detection]."""
        if self.model is None:
            raise RuntimeError("Model not loaded. Call load_model() first.")
        try:
            return self.detect_batch([frame])[0]
        except Exception as e:
            logger.error(f"Detection failed: {e}")
            raise

    def detect_batch(self, frames: Sequence[np.ndarray], strict: bool = False) -> list[list[Detection]]:
        """Batched detection.  Frames of equal size run as one device batch (chunks of `batch_size`).  A call that mixes frame
        sizes runs, like the reference's _preprocess_batch (DetrImageProcessor pads to the batch maximum and returns pixel_mask), as
        padded device batches with the mask carried through the transformer (`mixed_sizes="pad"`, the default), or as one device
        batch per frame size (`mixed_sizes="group"`: no padding arithmetic at all; also what detect_batch_with_features uses).

        Failure isolation (the reference's DetectionPhase catches exceptions PER FRAME and records an empty list,
        src/pipeline/phases/detection.py:124-127): a frame that is not a uint8 [H,W,3] array, or whose device batch fails and
        which then fails again on its own, yields [] and an error log - its batch neighbours keep their detections.
        `strict=True` raises instead."""
        return self._detect_batch(frames, with_features=False, strict=strict)[0]

    def _staging(self, n: int, h: int, w: int, slot: int = 0):
        """Pinned host staging buffer [batch_size, h, w, 3] (two per frame size - device batches are pipelined -, reused across
        calls)."""
        torch = _lib.require_cuda()
        buf = self._pinned.get((h, w, slot))
        if buf is None or buf.shape[0] < n:
            buf = torch.empty((max(n, min(self.batch_size, 64)), h, w, 3), dtype=torch.uint8).pin_memory()
            self._pinned[(h, w, slot)] = buf
        return buf

    def _stage(self, frames, chunk, h0, w0, slot: int = 0):
        """frames[chunk] -> the pinned staging buffer of their size (one copy per frame; large copies release the GIL, so a few
        threads share the 3 MB memcpys: 24 -> 6 ms for 64 frames of 800x1333)."""
        stage = self._staging(len(chunk), h0, w0, slot)
        view = stage[:len(chunk)]
        host = view.numpy()

        def put(j):
            host[j] = frames[chunk[j]]

        if len(chunk) >= 8:
            if self._copy_pool is None:
                from concurrent.futures import ThreadPoolExecutor

                self._copy_pool = ThreadPoolExecutor(max_workers=8, thread_name_prefix="opd-stage")
            list(self._copy_pool.map(put, range(len(chunk))))
        else:
            for j in range(len(chunk)):
                put(j)
        return view

    def _detect_chunk(self, frames, chunk, h0, w0, with_features, dev):
        """One device batch of equal-size frames -> ([list[Detection]] per frame, [features] per frame)."""
        return self._finish_chunk(self._launch_chunk(frames, chunk, h0, w0, with_features, dev, 0))

    def _launch_chunk(self, frames, chunk, h0, w0, with_features, dev, slot: int):
        """Stage the frames, enqueue the copy, the forward and the post-processing, and the device -> host copy of the packed result
        rows into pinned memory; returns without waiting (the caller stages the next device batch meanwhile)."""
        torch = _lib.require_cuda()
        view = self._stage(frames, chunk, h0, w0, slot)
        out = self.detect_tensors(view.to(dev, non_blocking=True))
        f_dev = self.model.roi_features(out["det_xywh"], out["n_keep"], h0, w0) if with_features else None
        # one packed read of the result rows: [B, Q, 4 + 1 + 2 + 1] float64
        packed = torch.cat([out["det_xywh"], out["det_score"].double().unsqueeze(-1), out["det_foot"],
                            out["det_query"].double().unsqueeze(-1)], dim=-1)
        host = {"rows": torch.empty(packed.shape, dtype=packed.dtype).pin_memory().copy_(packed, non_blocking=True),
                "n_keep": torch.empty(out["n_keep"].shape, dtype=out["n_keep"].dtype).pin_memory().copy_(out["n_keep"], non_blocking=True)}
        if with_features:
            host["feats"] = torch.empty(f_dev.shape, dtype=f_dev.dtype).pin_memory().copy_(f_dev, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        return host, done, len(chunk)

    @staticmethod
    def _finish_chunk(pending):
        host, done, n_frames = pending
        done.synchronize()
        return ViTDetector._rows_to_detections(host["rows"].numpy(), host["n_keep"].numpy(),
                                               host["feats"].numpy() if "feats" in host else None, n_frames)

    @staticmethod
    def _to_detections(out: dict, f_dev, n_frames: int):
        """Compacted device rows -> ([list[Detection]] per frame, [features] per frame)."""
        torch = _lib.require_cuda()
        packed = torch.cat([out["det_xywh"], out["det_score"].double().unsqueeze(-1), out["det_foot"],
                            out["det_query"].double().unsqueeze(-1)], dim=-1)
        return ViTDetector._rows_to_detections(packed.cpu().numpy(), out["n_keep"].cpu().numpy(),
                                               f_dev.cpu().numpy() if f_dev is not None else None, n_frames)

    @staticmethod
    def _rows_to_detections(rows, n_keep, f_host, n_frames: int):
        with_features = f_host is not None
        dets_all, feats_all = [], []
        for j in range(n_frames):
            n = int(n_keep[j])
            r = rows[j, :n].tolist()
            dets = [Detection(bbox=(x, y, w, h), confidence=float(np.float32(sc)), class_id=PERSON_LABEL, class_name="person",
                              camera_coords=(fx, fy), query_index=int(q)) for x, y, w, h, sc, fx, fy, q in r]
            feat = None
            if with_features:
                feat = f_host[j, :n].copy()
                for k, d in enumerate(dets):      # like the YOLO twin (yolov8_detector.py:153-156)
                    d.features = feat[k]
            dets_all.append(dets)
            feats_all.append(feat)
        return dets_all, feats_all

    def _detect_batch(self, frames: Sequence[np.ndarray], with_features: bool, strict: bool = False):
        if self.model is None:
            raise RuntimeError("Model not loaded. Call load_model() first.")
        torch = _lib.require_cuda()
        results: list[list[Detection] | None] = [None] * len(frames)
        feats: list[np.ndarray | None] = [None] * len(frames)
        groups: dict[tuple[int, int], list[int]] = {}
        for i, f in enumerate(frames):
            try:
                f = np.asarray(f)
                if f.ndim != 3 or f.shape[2] != 3 or f.dtype != np.uint8 or f.shape[0] < 1 or f.shape[1] < 1:
                    raise ValueError(f"frame {i}: expected a uint8 [H,W,3] BGR array, got {f.dtype} {f.shape}")
            except Exception as e:
                if strict:
                    raise
                logger.error(f"Detection failed for frame {i}: {e}")
                results[i] = []
                continue
            groups.setdefault((f.shape[0], f.shape[1]), []).append(i)
        dev = torch.device("cuda", self._device_index())
        if len(groups) > 1 and self.mixed_sizes == "pad" and not with_features:
            self._detect_padded(frames, groups, results, dev, strict)
            groups = {}
        # Device batches are pipelined: while batch k runs on the GPU the host stages batch k + 1 into the other pinned buffer and
        # then builds the Detection objects of batch k - 1.  A call that fits one device batch is split in two halves for that.
        per = self.batch_size
        total = sum(len(v) for v in groups.values())
        if total > 16 and all(len(v) <= per for v in groups.values()):
            per = max(8, (max(len(v) for v in groups.values()) + 1) // 2)
        chunks = [(h0, w0, idxs[c0:c0 + per]) for (h0, w0), idxs in groups.items() for c0 in range(0, len(idxs), per)]

        def store(chunk, d, f):
            for j, i in enumerate(chunk):
                results[i] = d[j]
                feats[i] = f[j]

        def careful(h0, w0, chunk):
            """A device batch that failed as a whole (e.g. frames too small for the backbone) is retried frame by frame, so that
            one bad frame costs only itself."""
            d, f = [], []
            for i in chunk:
                try:
                    di, fi = self._detect_chunk(frames, [i], h0, w0, with_features, dev)
                except _lib.OpdError as e1:
                    if strict:
                        raise
                    logger.error(f"Detection failed for frame {i}: {e1}")
                    di, fi = [[]], [None]
                d += di
                f += fi
            store(chunk, d, f)

        pending = None
        for k, (h0, w0, chunk) in enumerate(chunks):
            try:
                launched = self._launch_chunk(frames, chunk, h0, w0, with_features, dev, k % 2)
            except _lib.OpdError as e:
                if strict:
                    raise
                logger.error(f"Detection failed for a batch of {len(chunk)} frames ({e}); retrying frame by frame")
                launched = None
            if pending is not None:
                store(pending[0], *self._finish_chunk(pending[1]))
                pending = None
            if launched is None:
                careful(h0, w0, chunk)
            else:
                pending = (chunk, launched)
        if pending is not None:
            store(pending[0], *self._finish_chunk(pending[1]))
        return [r if r is not None else [] for r in results], feats

    def _detect_padded(self, frames, groups: dict, results: list, dev, strict: bool) -> None:
        """Mixed frame sizes as padded device batches (chunks of `batch_size` frames in call order)."""
        torch = _lib.require_cuda()
        order = sorted(i for idxs in groups.values() for i in idxs)
        for c0 in range(0, len(order), self.batch_size):
            chunk = order[c0:c0 + self.batch_size]
            by_size: dict[tuple[int, int], list[int]] = {}
            for i in chunk:
                by_size.setdefault((frames[i].shape[0], frames[i].shape[1]), []).append(i)
            try:
                tensors = []
                for (h0, w0), idxs in by_size.items():
                    tensors.append(self._stage(frames, idxs, h0, w0).to(dev, non_blocking=True))
                logits, boxes = self.model.forward_mixed(tensors) if len(tensors) > 1 else self.model.forward(tensors[0])
                b0 = 0
                for (h0, w0), idxs in by_size.items():
                    n = len(idxs)
                    out = postprocess_tensors(logits[b0:b0 + n].contiguous(), boxes[b0:b0 + n].contiguous(), h0, w0,
                                              self.confidence_threshold)
                    dets, _ = self._to_detections(out, None, n)
                    for j, i in enumerate(idxs):
                        results[i] = dets[j]
                    b0 += n
            except _lib.OpdError as e:
                if strict:
                    raise
                logger.error(f"Detection failed for a padded batch of {len(chunk)} frames: {e}")
                for i in chunk:
                    results[i] = []

    def _get_foot_position(self, bbox: tuple[float, float, float, float]) -> tuple[float, float]:
        x, y, w, h = bbox
        return (x + w / 2, y + h)

    def detect_with_features(self, frame: np.ndarray) -> tuple[list[Detection], np.ndarray]:
        """(detections, features [n, 256] float32): the call DetectionPhase.execute makes per frame
        (src/pipeline/phases/detection.py:94).  Features = ROI mean-pool of the DETR encoder output over each box,
        L2-normalised (the removed ViTDetector._extract_features_from_outputs; arithmetic of
        src/tracking/feature_extractor.py:39-88), computed on the device from the same forward."""
        dets, feats = self.detect_batch_with_features([frame])
        return dets[0], feats[0]

    def detect_batch_with_features(self, frames: Sequence[np.ndarray]) -> tuple[list[list[Detection]], list[np.ndarray]]:
        """Batched detect_with_features: one forward per device batch, features for every kept box."""
        dets, feats = self._detect_batch(frames, with_features=True)
        return dets, [f if f is not None else np.zeros((0, 256), np.float32) for f in feats]

    def extract_features(self, frame: np.ndarray, detections: list[Detection]) -> np.ndarray:
        """Encoder ROI features [len(detections), 256] float32 of GIVEN boxes on `frame` (one forward of the frame)."""
        if self.model is None:
            raise RuntimeError("Model not loaded. Call load_model() first.")
        if len(detections) == 0:
            return np.zeros((0, 256), np.float32)
        torch = _lib.require_cuda()
        f = np.ascontiguousarray(frame)
        dev = torch.device("cuda", self._device_index())
        self.forward_raw(torch.from_numpy(f[None]).to(dev))
        out = []
        for c0 in range(0, len(detections), N_QUERIES):      # rows of one [1, Q, 4] box table
            chunk = detections[c0:c0 + N_QUERIES]
            xywh = torch.zeros(1, N_QUERIES, 4, dtype=torch.float64)
            xywh[0, :len(chunk)] = torch.tensor([list(d.bbox) for d in chunk], dtype=torch.float64)
            n = torch.tensor([len(chunk)], dtype=torch.int32)
            out.append(self.model.roi_features(xywh.to(dev), n.to(dev), f.shape[0], f.shape[1])[0, :len(chunk)].cpu().numpy())
        return np.concatenate(out, axis=0)

    def get_attention_map(self, _frame: np.ndarray, _layer_index: int = -1):
        logger.warning("Attention maps are a debug aid of the reference and are not produced by the fused attention kernel")
        return None
