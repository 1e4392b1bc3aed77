"""Synthetic DETR-ResNet-50 weights and frames (data generators only, no model arithmetic).

There is no network in the build / benchmark environment, so `facebook/detr-resnet-50` cannot be downloaded:
benchmarks and parity tests use a seeded random-init state dict of the same architecture (transformers key
names, float32).  Both sides of every parity test (this package and oracle/detr_oracle.py / transformers' own
DetrForObjectDetection) load the SAME dict."""

from __future__ import annotations

import math

import numpy as np
import torch

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


def conv_specs():
    """(hf_prefix, c_in, c_out, k, stride) of every backbone convolution, in execution order."""
    specs = [("model.backbone.model.embedder.embedder", 3, EMBED, 7, 2)]
    c_in = EMBED
    for s, (depth, width) in enumerate(zip(STAGE_DEPTHS, STAGE_WIDTHS)):
        mid = width // 4
        for l in range(depth):
            stride = 2 if (l == 0 and s > 0) else 1
            p = f"model.backbone.model.encoder.stages.{s}.layers.{l}"
            if l == 0:
                specs.append((p + ".shortcut", c_in, width, 1, stride))
            specs.append((p + ".layer.0", c_in, mid, 1, 1))
            specs.append((p + ".layer.1", mid, mid, 3, stride))
            specs.append((p + ".layer.2", mid, width, 1, 1))
            c_in = width
    return specs


def random_init_state_dict(seed: int = 0, trained_like: bool = False) -> dict[str, torch.Tensor]:
    """Seeded random-init DETR-R50 state dict (transformers key names, float32).

    `trained_like=True` is the variance-preserving set used by the parity study: every attention
    projection at Xavier gain 1 (no sharpened softmax, no damped o_proj), LayerNorm weights at 1 +- 0.05, the other tensors as
    below - per-layer gain ~ 1, so that a rounding difference is carried through the ~100 layers instead of being amplified.

    transformers' default init is degenerate for parity purposes (every query yields the same box, no score
    crosses 0.5 — SURVEY.md H2), so the variances are re-scaled: He-init convolutions with non-trivial frozen-BN
    statistics, unit-scale attention / FFN weights, wide query embeddings and heads.  Both sides of every
    parity test load this same dict.
    """
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    def rand(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    w: dict[str, torch.Tensor] = {}
    for prefix, c_in, c_out, k, _stride in conv_specs():
        fan_in = c_in * k * k
        w[prefix + ".convolution.weight"] = randn(c_out, c_in, k, k, std=math.sqrt(2.0 / fan_in))
        last_of_block = prefix.endswith(".layer.2")
        scale = 0.35 if last_of_block else 1.0          # keep the residual stream from blowing up
        w[prefix + ".normalization.weight"] = rand(c_out, lo=0.6, hi=1.4) * scale
        w[prefix + ".normalization.bias"] = randn(c_out, std=0.1)
        w[prefix + ".normalization.running_mean"] = randn(c_out, std=0.1)
        w[prefix + ".normalization.running_var"] = rand(c_out, lo=0.6, hi=1.4)

    w["model.input_projection.weight"] = randn(D_MODEL, STAGE_WIDTHS[-1], 1, 1, std=0.2 / math.sqrt(STAGE_WIDTHS[-1]))
    w["model.input_projection.bias"] = randn(D_MODEL, std=0.1)
    w["model.query_position_embeddings.weight"] = randn(N_QUERIES, D_MODEL, std=1.0)

    def linear(prefix, n_out, n_in, std=None, bias_std=0.05):
        w[prefix + ".weight"] = randn(n_out, n_in, std=std if std is not None else 1.0 / math.sqrt(n_in))
        w[prefix + ".bias"] = randn(n_out, std=bias_std)

    def layer_norm(prefix):
        w[prefix + ".weight"] = rand(D_MODEL, lo=0.95, hi=1.05) if trained_like else rand(D_MODEL, lo=0.8, hi=1.2)
        w[prefix + ".bias"] = randn(D_MODEL, std=0.05)

    def attn(prefix, qk_gain=1.0, o_gain=1.0):
        if trained_like:
            qk_gain = o_gain = 1.0
        # qk_gain > 1 sharpens the softmax so that different queries attend to different tokens; o_gain < 1 keeps
        # the residual stream token-specific (random post-norm attention stacks otherwise collapse to one token)
        for proj in ("q_proj", "k_proj", "v_proj", "o_proj"):
            gain = qk_gain if proj in ("q_proj", "k_proj") else (o_gain if proj == "o_proj" else 1.0)
            linear(f"{prefix}.{proj}", D_MODEL, D_MODEL, std=gain / math.sqrt(D_MODEL))

    for i in range(N_ENC):
        p = f"model.encoder.layers.{i}"
        attn(p + ".self_attn", qk_gain=1.5, o_gain=0.3)
        layer_norm(p + ".self_attn_layer_norm")
        linear(p + ".mlp.fc1", FFN, D_MODEL)
        linear(p + ".mlp.fc2", D_MODEL, FFN)
        layer_norm(p + ".final_layer_norm")
    for i in range(N_DEC):
        p = f"model.decoder.layers.{i}"
        attn(p + ".self_attn", qk_gain=1.5, o_gain=0.3)
        layer_norm(p + ".self_attn_layer_norm")
        attn(p + ".encoder_attn", qk_gain=3.0)
        layer_norm(p + ".encoder_attn_layer_norm")
        linear(p + ".mlp.fc1", FFN, D_MODEL)
        linear(p + ".mlp.fc2", D_MODEL, FFN)
        layer_norm(p + ".final_layer_norm")
    layer_norm("model.decoder.layernorm")
    linear("class_labels_classifier", N_CLASSES + 1, D_MODEL, std=0.2, bias_std=0.3)
    # make "person" competitive so that a useful fraction of queries is a person above the usual thresholds
    w["class_labels_classifier.bias"][PERSON_LABEL] += 13.4   # calibrated on synthetic_frames: person ~ the dominant class
    linear("bbox_predictor.layers.0", D_MODEL, D_MODEL)
    linear("bbox_predictor.layers.1", D_MODEL, D_MODEL)
    linear("bbox_predictor.layers.2", 4, D_MODEL, std=0.12, bias_std=0.3)
    return w



def synthetic_frames(batch: int, h: int, w: int, seed: int = 1) -> np.ndarray:
    """Synthetic BGR frames, uint8: large flat-colour rectangles (so that distant image regions give distinct
    backbone features) with finer blocks and pixel noise on top."""
    rng = np.random.default_rng(seed)

    def blocks(size, amp):
        c = rng.integers(-amp, amp + 1, (batch, (h + size - 1) // size, (w + size - 1) // size, 3), dtype=np.int16)
        return np.repeat(np.repeat(c, size, axis=1), size, axis=2)[:, :h, :w]

    img = 128 + blocks(192, 110) + blocks(32, 40) + rng.integers(-12, 13, (batch, h, w, 3), dtype=np.int16)
    return np.clip(img, 0, 255).astype(np.uint8)


def device_frames_reference(seed_base: int, global_frame0: int, batch: int, h: int, w: int) -> np.ndarray:
    """NumPy restatement of opd_synthetic_frames_u8 (csrc/detr_kernels.cu synthetic_frames_kernel): the frames bench.py's
    config 4 generates on the device, bit for bit."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)

    def mix(z):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return ((z ^ (z >> np.uint64(31))) >> np.uint64(32)).astype(np.int64)

    out = np.empty((batch, h, w, 3), np.uint8)
    yy, xx, cc = np.meshgrid(np.arange(h, dtype=np.uint64), np.arange(w, dtype=np.uint64), np.arange(3, dtype=np.uint64), indexing="ij")
    with np.errstate(over="ignore"):
        for b in range(batch):
            g = global_frame0 + b
            base = (np.uint64((seed_base + g // 64) & int(M)) * np.uint64(0x9E3779B97F4A7C15) +
                    np.uint64(g % 64) * np.uint64(0xD1B54A32D192ED03))
            v = np.full((h, w, 3), 128, np.int64)
            for level, div, mod, off in ((0, 192, 221, 110), (1, 32, 81, 40), (2, 1, 25, 12)):
                key = ((np.uint64(level * 4096) + yy // np.uint64(div)) * np.uint64(4096) + xx // np.uint64(div)) * np.uint64(4) + cc
                v += mix(base + key * np.uint64(0x94D049BB133111EB)) % mod - off
            out[b] = np.clip(v, 0, 255).astype(np.uint8)
    return out
