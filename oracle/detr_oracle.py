"""CPU oracle of the DETR-ResNet-50 person detector (Phase 2 of the hot path).  TEST INFRASTRUCTURE ONLY:
nothing under office_person_detection_vit_b200/ may import this module; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs use it, as the checker or the timed CPU baseline.

The reference's detector file (src/detection/vit_detector.py) is absent from the snapshot;
its arithmetic lives in the third-party dependency `transformers` (pinned 4.57.3 in the reference's
requirements.txt:220; 5.5.0 is what this image has, same architecture).  This module restates that published
algorithm with plain torch functional ops, citing the transformers file:line each function follows
(paths relative to site-packages/transformers/):

    preprocess      models/detr/image_processing_detr.py:687-801, image_processing_backends.py:200-252,308-331,
                    image_transforms.py:206-242
    frozen BN       models/detr/modeling_detr.py:185-222
    ResNet-50       models/resnet/modeling_resnet.py:40-240  (v1.5: stride on the 3x3, downsample_in_bottleneck=False)
    sine pos-embed  models/detr/modeling_detr.py:300-349
    attention       models/detr/modeling_detr.py:386-557
    enc/dec layers  models/detr/modeling_detr.py:560-723, 929-1100
    model forward   models/detr/modeling_detr.py:1142-1272
    heads           models/detr/modeling_detr.py:1275-1297, 1401-1402
    postprocess     models/detr/image_processing_detr.py:804-855, image_transforms.py:529-536
    person filter + xyxy->xywh + foot point: the reference's surviving detector
                    src/detection/yolov8_detector.py:210-225, 229-241

Pinning: tests/test_detr_oracle.py checks `forward(mode="fp32")` against transformers' own
DetrForObjectDetection + DetrImageProcessor on identical weights and frames (that IS the reference's
arithmetic), and against the golden vectors committed under tests/golden/ (made by tests/golden/make_detr_golden.py).

Two arithmetic modes:
  "fp32"  the reference arithmetic (float32 everywhere);
  "bf16"  float32 accumulation with the SAME bfloat16 rounding points as the CUDA path:
          BN folded into the conv weights in float32 and then rounded to bf16, every stored activation rounded
          to bf16, softmax / LayerNorm statistics / heads in float32.
Study modes (oracle/parity_study.py: where does the bf16 error come from, what would a mixed-precision path buy):
  "bf16+res32"  as "bf16", but the transformer's residual stream / LayerNorm inputs and outputs stay float32 (only the GEMM
                and attention operands are rounded);
  "bf16bb"      bf16 backbone + input projection, float32 transformer;   "bf16tr"  float32 backbone, bf16 transformer.
"""

from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# data generators shared by both sides of every parity test (no arithmetic of the model lives there)
from office_person_detection_vit_b200.detection.synthetic import (  # noqa: E402,F401
    conv_specs, random_init_state_dict as make_weights, synthetic_frames)

STAGE_DEPTHS = (3, 4, 6, 3)
STAGE_WIDTHS = (256, 512, 1024, 2048)
EMBED = 64
D_MODEL = 256
N_HEADS = 8
FFN = 2048
N_ENC = 6
N_DEC = 6
N_QUERIES = 100
N_CLASSES = 91           # logits have N_CLASSES + 1 entries (last = "no object")
PERSON_LABEL = 1         # COCO id of "person" in facebook/detr-resnet-50
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)
BN_EPS = 1e-5
LN_EPS = 1e-5


# ------------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------------
def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


MODES = ("fp32", "bf16", "bf16+res32", "bf16bb", "bf16tr")


class _Mode:
    def __init__(self, mode: str, part: str = "transformer"):
        if mode not in MODES:
            raise ValueError(mode)
        self.bf16 = mode in ("bf16", "bf16+res32") or (mode == "bf16bb" and part == "backbone") or \
            (mode == "bf16tr" and part == "transformer")
        self.res32 = mode == "bf16+res32"

    def act(self, x):      # a stored activation / a tensor-core activation operand
        return _bf16(x) if self.bf16 else x

    def stream(self, x):   # the transformer's residual stream (LayerNorm output that the next residual add reads)
        return x if self.res32 else self.act(x)

    def wt(self, x):       # a tensor-core weight operand
        return _bf16(x) if self.bf16 else x


def resized_size(h: int, w: int, size: int = 800, max_size: int = 1333) -> tuple[int, int]:
    """image_transforms.py:206-242 get_size_with_aspect_ratio."""
    raw = None
    mn, mx = float(min(h, w)), float(max(h, w))
    if mx / mn * size > max_size:
        raw = max_size * mn / mx
        size = int(round(raw))
    if (h <= w and h == size) or (w <= h and w == size):
        return h, w
    if w < h:
        return (int(raw * h / w) if raw is not None else int(size * h / w)), size
    return size, (int(raw * w / h) if raw is not None else int(size * w / h))


def _aa_table(in_size: int, out_size: int):
    """ATen aten/src/ATen/native/cpu/UpSampleKernel.cpp (_compute_indices_min_size_weights_aa,
    _compute_index_ranges_int16_weights), bilinear: first source index, int16 weights [out, ksize], precision."""
    scale = in_size / out_size
    support = scale if scale >= 1.0 else 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    x0 = np.zeros(out_size, np.int64)
    wt = np.zeros((out_size, ksize), np.float64)
    wt_max = 0.0
    for i in range(out_size):
        center = scale * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xsize = min(max(min(int(center + support + 0.5), in_size) - xmin, 0), ksize)
        w = np.array([max(0.0, 1.0 - abs((j + xmin - center + 0.5) * invscale)) for j in range(xsize)], np.float64)
        total = w.sum() if xsize else 0.0
        if total != 0.0:
            w = w / total
            wt_max = max(wt_max, float(w.max()))
        wt[i, :xsize] = w
        x0[i] = xmin
    precision = 0
    while precision < 22 and int(0.5 + wt_max * (1 << (precision + 1))) < (1 << 15):
        precision += 1
    v = wt * (1 << precision)
    w16 = np.where(v < 0, (-0.5 + v).astype(np.int64), (0.5 + v).astype(np.int64))
    return x0, w16, ksize, precision


def resize_u8_antialias(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """[..., H, W] uint8 -> [..., out_h, out_w] uint8: the separable fixed-point kernel torch runs for
    F.interpolate(uint8, mode="bilinear", antialias=True) on the CPU (horizontal pass, uint8 intermediate, vertical
    pass).  Independent numpy restatement used to pin the CUDA resize; checked bit-exact against torch in tests."""
    def one_axis(a, axis, out_size):
        in_size = a.shape[axis]
        if in_size == out_size:
            return a
        x0, w16, ksize, prec = _aa_table(in_size, out_size)
        src = np.moveaxis(a, axis, 0).astype(np.int64)
        idx = np.minimum(x0[:, None] + np.arange(ksize)[None, :], in_size - 1)          # [out, k]
        acc = np.full((out_size,) + src.shape[1:], 1 << (prec - 1), np.int64)
        for j in range(ksize):
            acc += src[idx[:, j]] * w16[:, j].reshape((-1,) + (1,) * (src.ndim - 1))
        return np.moveaxis(np.clip(acc >> prec, 0, 255).astype(np.uint8), 0, axis)

    return one_axis(one_axis(np.asarray(img), -1, out_w), -2, out_h)


def preprocess(frames_bgr: np.ndarray | torch.Tensor, do_resize: bool = True) -> torch.Tensor:
    """[B,H0,W0,3] uint8 BGR -> pixel_values [B,3,H,W] float32 (all frames the same size: no padding, mask = 1).

    BGR->RGB (removed ViTDetector._preprocess, coverage.json lines 285-300), uint8 bilinear antialias resize
    (image_processing_backends.py:200-252 -> torchvision resize on the uint8 tensor), then the fused
    rescale+normalise (x - 255*mean) / (255*std) in float32 (image_processing_backends.py:308-331).
    """
    x = torch.as_tensor(np.ascontiguousarray(frames_bgr)) if not torch.is_tensor(frames_bgr) else frames_bgr
    assert x.dtype == torch.uint8 and x.ndim == 4 and x.shape[-1] == 3
    x = x.flip(-1).permute(0, 3, 1, 2).contiguous()          # RGB, CHW
    h0, w0 = x.shape[-2:]
    h, w = resized_size(h0, w0) if do_resize else (h0, w0)
    if (h, w) != (h0, w0):
        x = F.interpolate(x, size=(h, w), mode="bilinear", antialias=True, align_corners=False)
    mean = torch.tensor(IMAGE_MEAN) * 255.0       # float32, like transformers (tensor(mean) * (1 / rescale_factor))
    std = torch.tensor(IMAGE_STD) * 255.0
    return (x.to(torch.float32) - mean[None, :, None, None]) / std[None, :, None, None]


def fold_bn(w: dict, prefix: str) -> tuple[torch.Tensor, torch.Tensor]:
    """modeling_detr.py:211-222: y = conv(x) * scale + (bias - mean * scale), scale = weight * rsqrt(var + eps)."""
    scale = w[prefix + ".normalization.weight"] * (w[prefix + ".normalization.running_var"] + BN_EPS).rsqrt()
    shift = w[prefix + ".normalization.bias"] - w[prefix + ".normalization.running_mean"] * scale
    return w[prefix + ".convolution.weight"] * scale[:, None, None, None], shift


def sine_position_embedding(h: int, w: int) -> torch.Tensor:
    """modeling_detr.py:322-349 with an all-ones mask, num_position_features=128, normalize=True -> [h*w, 256]."""
    npf = D_MODEL // 2
    y_embed = torch.arange(1, h + 1, dtype=torch.float32)[:, None].expand(h, w)
    x_embed = torch.arange(1, w + 1, dtype=torch.float32)[None, :].expand(h, w)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[-1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, -1:] + eps) * scale
    dim_t = torch.arange(npf, dtype=torch.int64).to(torch.float32)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
    pos_x = x_embed[:, :, None] / dim_t
    pos_y = y_embed[:, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, 0::2].sin(), pos_x[:, :, 1::2].cos()), dim=3).flatten(2)
    pos_y = torch.stack((pos_y[:, :, 0::2].sin(), pos_y[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((pos_y, pos_x), dim=2).reshape(h * w, D_MODEL)


def sine_position_embedding_masked(mask: torch.Tensor) -> torch.Tensor:
    """modeling_detr.py:322-349 with a real pixel mask [B, h, w] (bool, True = picture) -> [B, h*w, 256]."""
    npf = D_MODEL // 2
    y_embed = mask.cumsum(1, dtype=torch.float32)
    x_embed = mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(npf, dtype=torch.int64).to(torch.float32)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).flatten(1, 2)


def _linear(m: _Mode, w: dict, prefix: str, x: torch.Tensor) -> torch.Tensor:
    # the activation operand of a tensor-core GEMM is bf16 (a no-op in plain "bf16" mode, where every input already is)
    return F.linear(m.act(x), m.wt(w[prefix + ".weight"]), w[prefix + ".bias"])


def _mha(m: _Mode, w: dict, prefix: str, q_in, k_in, v_in, key_mask=None) -> torch.Tensor:
    """modeling_detr.py:386-411, 441-477, 508-557: softmax(QK^T / sqrt(d)) V per head; key_mask [B, Lk] (bool, True = takes part):
    the key-padding mask of a padded batch - masked keys get probability 0."""
    B, Lq, _ = q_in.shape
    Lk = k_in.shape[1]
    dh = D_MODEL // N_HEADS
    q = m.act(_linear(m, w, prefix + ".q_proj", q_in)).view(B, Lq, N_HEADS, dh).transpose(1, 2)
    k = m.act(_linear(m, w, prefix + ".k_proj", k_in)).view(B, Lk, N_HEADS, dh).transpose(1, 2)
    v = m.act(_linear(m, w, prefix + ".v_proj", v_in)).view(B, Lk, N_HEADS, dh).transpose(1, 2)
    s = torch.matmul(q, k.transpose(2, 3)) * dh ** -0.5
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))
    if m.bf16:
        # CUDA path: P is rounded to bf16 for the PV product, the row sum stays float32 (flash-attention style)
        s = s - s.amax(dim=-1, keepdim=True)
        p = s.exp()
        o = torch.matmul(_bf16(p), v) / p.sum(dim=-1, keepdim=True)
    else:
        o = torch.matmul(F.softmax(s, dim=-1), v)
    o = m.act(o.transpose(1, 2).reshape(B, Lq, D_MODEL))
    return _linear(m, w, prefix + ".o_proj", o)


def _ln(w: dict, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (D_MODEL,), w[prefix + ".weight"], w[prefix + ".bias"], LN_EPS)


# ------------------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def backbone(w: dict, pixel_values: torch.Tensor, mode: str = "fp32", taps: dict | None = None) -> torch.Tensor:
    """ResNet-50 with frozen BN -> stage-4 feature map [B,2048,h,w]."""
    m = _Mode(mode, "backbone")

    def conv(prefix, x, stride, k, relu, residual=None, rounded=True):
        wt, shift = fold_bn(w, prefix)
        y = F.conv2d(x, m.wt(wt), shift, stride=stride, padding=k // 2)
        if residual is not None:
            y = y + residual
        if relu:
            y = F.relu(y)
        return m.act(y) if rounded else y

    x = m.act(pixel_values)
    x = conv("model.backbone.model.embedder.embedder", x, 2, 7, True)
    if taps is not None:
        taps["stem"] = x
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps["pool"] = x
    for s, depth in enumerate(STAGE_DEPTHS):
        for l in range(depth):
            p = f"model.backbone.model.encoder.stages.{s}.layers.{l}"
            stride = 2 if (l == 0 and s > 0) else 1
            # CUDA path: the projection shortcut of stages 1, 3 and 4 is accumulated in fp32 together with layer.2 (second
            # k-range of the same GEMM) and never stored; stage 2's is a stored bf16 tensor
            res = conv(p + ".shortcut", x, stride, 1, False, rounded=(s == 1)) if l == 0 else x
            y = conv(p + ".layer.0", x, 1, 1, True)
            y = conv(p + ".layer.1", y, stride, 3, True)
            x = conv(p + ".layer.2", y, 1, 1, True, residual=res)
            if taps is not None:
                taps[f"stage{s}.{l}"] = x
    return x


def _transformer(w: dict, mode: str, feat: torch.Tensor, pos: torch.Tensor, key_mask, taps: dict | None):
    """input projection, 6 encoder + 6 decoder layers, heads.  feat [B,2048,h,w]; pos [1 or B, h*w, 256]; key_mask [B, h*w] or None."""
    m = _Mode(mode)
    mb = _Mode(mode, "backbone")
    B = feat.shape[0]
    proj = F.conv2d(feat, mb.wt(w["model.input_projection.weight"]), w["model.input_projection.bias"])
    x = m.stream(mb.act(proj.flatten(2).permute(0, 2, 1)) if not m.res32 else proj.flatten(2).permute(0, 2, 1))   # [B, S, 256]
    if taps is not None:
        taps["enc_in"] = x
        taps["pos"] = pos.reshape(-1, D_MODEL)

    for i in range(N_ENC):
        p = f"model.encoder.layers.{i}"
        qk = m.act(x + pos)
        a = _mha(m, w, p + ".self_attn", qk, qk, x, key_mask)
        x = m.stream(_ln(w, p + ".self_attn_layer_norm", x + a))
        f = m.act(F.relu(_linear(m, w, p + ".mlp.fc1", x)))
        f = _linear(m, w, p + ".mlp.fc2", f)
        x = m.stream(_ln(w, p + ".final_layer_norm", x + f))
        if taps is not None:
            taps[f"enc{i}"] = x
    memory = x
    mem_k = m.act(memory + pos)

    qpos = w["model.query_position_embeddings.weight"][None].expand(B, -1, -1)
    y = torch.zeros(B, N_QUERIES, D_MODEL)
    for i in range(N_DEC):
        p = f"model.decoder.layers.{i}"
        qk = m.act(y + qpos)
        a = _mha(m, w, p + ".self_attn", qk, qk, y)
        y = m.stream(_ln(w, p + ".self_attn_layer_norm", y + a))
        a = _mha(m, w, p + ".encoder_attn", m.act(y + qpos), mem_k, memory, key_mask)
        y = m.stream(_ln(w, p + ".encoder_attn_layer_norm", y + a))
        f = m.act(F.relu(_linear(m, w, p + ".mlp.fc1", y)))
        f = _linear(m, w, p + ".mlp.fc2", f)
        y = m.stream(_ln(w, p + ".final_layer_norm", y + f))
        if taps is not None:
            taps[f"dec{i}"] = y
    y = m.stream(_ln(w, "model.decoder.layernorm", y))
    if taps is not None:
        taps["dec_out"] = y

    # heads stay float32 (weights too) in both modes: modeling_detr.py:1275-1297, 1401-1402
    logits = F.linear(y, w["class_labels_classifier.weight"], w["class_labels_classifier.bias"])
    b = F.relu(F.linear(y, w["bbox_predictor.layers.0.weight"], w["bbox_predictor.layers.0.bias"]))
    b = F.relu(F.linear(b, w["bbox_predictor.layers.1.weight"], w["bbox_predictor.layers.1.bias"]))
    boxes = F.linear(b, w["bbox_predictor.layers.2.weight"], w["bbox_predictor.layers.2.bias"]).sigmoid()
    return logits, boxes


@torch.no_grad()
def forward(w: dict, frames_bgr, mode: str = "fp32", taps: dict | None = None, do_resize: bool = True):
    """frames [B,H0,W0,3] uint8 BGR -> (logits [B,100,92], boxes cxcywh in [0,1] [B,100,4]), float32."""
    pv = preprocess(frames_bgr, do_resize)
    if taps is not None:
        taps["pixel_values"] = pv
    feat = backbone(w, pv, mode, taps)
    pos = sine_position_embedding(feat.shape[2], feat.shape[3])[None]                        # float32 table
    return _transformer(w, mode, feat, pos, None, taps)


@torch.no_grad()
def forward_mixed(w: dict, frames_list, mode: str = "fp32", taps: dict | None = None, do_resize: bool = True):
    """A batch that mixes frame sizes: `frames_list` = list of [H0,W0,3] uint8 BGR frames.  Every frame is preprocessed on its own,
    zero-padded (after normalisation) at the bottom / right to the batch maximum with pixel_mask = 0 there
    (image_processing_detr.py:638-667), the mask is downsampled to the feature map by nearest interpolation
    (modeling_detr.py:281-283), drives the sine embedding (:322-349) and masks the keys of the encoder self-attention and the
    decoder cross-attention (:386-411)."""
    pvs = [preprocess(np.asarray(f)[None], do_resize) for f in frames_list]
    Hc, Wc = max(p.shape[2] for p in pvs), max(p.shape[3] for p in pvs)
    pv = torch.zeros(len(pvs), 3, Hc, Wc)
    mask = torch.zeros(len(pvs), Hc, Wc, dtype=torch.bool)
    for i, p_ in enumerate(pvs):
        pv[i, :, :p_.shape[2], :p_.shape[3]] = p_[0]
        mask[i, :p_.shape[2], :p_.shape[3]] = True
    if taps is not None:
        taps["pixel_values"] = pv
    feat = backbone(w, pv, mode, taps)
    fmask = F.interpolate(mask[None].float(), size=feat.shape[-2:]).to(torch.bool)[0]      # [B, h, w]
    pos = sine_position_embedding_masked(fmask)
    return _transformer(w, mode, feat, pos, fmask.flatten(1), taps)


@torch.no_grad()
def postprocess(logits: torch.Tensor, boxes: torch.Tensor, h0: int, w0: int):
    """image_processing_detr.py:826-843 without the threshold: per query (score, label, xyxy in pixels of the
    ORIGINAL frame)."""
    prob = F.softmax(logits, -1)
    scores, labels = prob[..., :-1].max(-1)
    cx, cy, bw, bh = boxes.unbind(-1)
    xyxy = torch.stack([cx - 0.5 * bw, cy - 0.5 * bh, cx + 0.5 * bw, cy + 0.5 * bh], dim=-1)
    xyxy = xyxy * torch.tensor([w0, h0, w0, h0], dtype=torch.float32)
    return scores, labels, xyxy


@torch.no_grad()
def detections(logits, boxes, h0: int, w0: int, threshold: float, person_label: int = PERSON_LABEL):
    """Per frame: list of (x, y, w, h, score, foot_x, foot_y) for score > threshold and label == person,
    in query order (post_process_object_detection + yolov8_detector.py:210-225, 229-241)."""
    scores, labels, xyxy = postprocess(logits, boxes, h0, w0)
    out = []
    for s, l, b in zip(scores, labels, xyxy):
        keep = (s > threshold) & (l == person_label)
        rows = []
        for sc, bb in zip(s[keep].tolist(), b[keep].tolist()):
            x1, y1, x2, y2 = bb
            bw, bh = x2 - x1, y2 - y1
            rows.append((x1, y1, bw, bh, sc, x1 + bw / 2, y1 + bh))
        out.append(rows)
    return out


def roi_features(encoder_features: np.ndarray, bboxes, image_shape: tuple[int, int]) -> np.ndarray:
    """ROI mean-pool + L2 normalisation of the encoder map (the removed ViTDetector.extract_features): restates
    FeatureExtractor.extract_roi_features (src/tracking/feature_extractor.py:39-88) and normalize_features (:21-37).
    encoder_features [h, w, D]; bboxes (x, y, width, height) in pixels of image_shape = (height, width)."""
    h, w, dim = encoder_features.shape
    img_h, img_w = image_shape
    rows = []
    for x, y, width, height in bboxes:
        x_min = int((x / img_w) * w)
        y_min = int((y / img_h) * h)
        x_max = int(((x + width) / img_w) * w)
        y_max = int(((y + height) / img_h) * h)
        x_min = max(0, min(x_min, w - 1))
        y_min = max(0, min(y_min, h - 1))
        x_max = max(x_min + 1, min(x_max, w))
        y_max = max(y_min + 1, min(y_max, h))
        rows.append(encoder_features[y_min:y_max, x_min:x_max, :].mean(axis=(0, 1)))
    if not rows:
        return np.array([]).reshape(0, dim)
    f = np.array(rows)
    return f / (np.linalg.norm(f, axis=1, keepdims=True) + 1e-8)


def hf_model(w: dict):
    """transformers' own DetrForObjectDetection loaded with `w` (the arithmetic the reference ran)."""
    import os

    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    from transformers import DetrConfig, DetrForObjectDetection, ResNetConfig

    cfg = DetrConfig(backbone_config=ResNetConfig(out_features=["stage4"]), num_labels=N_CLASSES)
    model = DetrForObjectDetection(cfg).eval()
    missing, unexpected = model.load_state_dict(w, strict=False)
    assert not unexpected and all("num_batches_tracked" in k for k in missing), (missing, unexpected)
    return model
