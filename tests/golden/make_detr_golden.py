"""Generates tests/golden/detr_small.npz: outputs of transformers' OWN DetrForObjectDetection + DetrImageProcessor
(the arithmetic the reference's removed ViTDetector drove) on seeded synthetic frames with the seeded random-init
weights.  Run here (CPU): python tests/golden/make_detr_golden.py"""

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("HF_HUB_OFFLINE", "1")

from oracle import detr_oracle as do  # noqa: E402


def main():
    from transformers import DetrImageProcessor

    w = do.make_weights(0)
    model = do.hf_model(w)
    out = {}
    # (a) un-resized small frames; (b) a 180x320 frame that the processor resizes (longest-edge clamp: 750x1333)
    for tag, frames, proc in (
            ("small", do.synthetic_frames(2, 96, 128, seed=7), DetrImageProcessor(do_resize=False)),
            ("resized", do.synthetic_frames(1, 180, 320, seed=8), DetrImageProcessor())):
        rgb = [np.ascontiguousarray(f[:, :, ::-1]) for f in frames]
        with torch.no_grad():
            inp = proc(images=rgb, return_tensors="pt")
            o = model(**inp)
        out[f"{tag}_frames"] = frames
        out[f"{tag}_pixel_values"] = inp["pixel_values"].numpy().astype(np.float32)
        out[f"{tag}_logits"] = o.logits.numpy()
        out[f"{tag}_boxes"] = o.pred_boxes.numpy()
    np.savez_compressed(Path(__file__).resolve().parent / "detr_small.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
