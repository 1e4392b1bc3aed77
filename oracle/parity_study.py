"""CPU study behind DESIGN.md "numerics" (VERDICT r1 item 1): how far do the bf16 rounding points alone move the reference
arithmetic, which part of the network does it, and what would keeping the transformer's residual stream in float32 buy?

TEST INFRASTRUCTURE (oracle side only, nothing here is product code).  Runs the oracle (oracle/detr_oracle.py) in its
float32 mode and in each bf16 variant on the same frames and weights and prints / stores, per variant: relative L2 error of
every tap against the float32 run, box error in pixels (max / median over all queries), score error, label agreement.

    python -m oracle.parity_study [--size 800 1333] [--frames 2] [--seeds 0 1] [--out profiles/r02_parity_oracle_study.json]
"""

from __future__ import annotations

import argparse
import json
import time

import torch

from oracle import detr_oracle as do


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def summarise(logits, boxes, ref_logits, ref_boxes, h0, w0):
    sc, lb, xy = do.postprocess(logits, boxes, h0, w0)
    rs, rl, rx = do.postprocess(ref_logits, ref_boxes, h0, w0)
    e = (xy - rx).abs()
    return {"box_px_max": float(e.max()), "box_px_median": float(e.median()), "box_px_p99": float(e.flatten().quantile(0.99)),
            "score_max": float((sc - rs).abs().max()), "score_median": float((sc - rs).abs().median()),
            "label_agreement": float((lb == rl).float().mean()), "logits_rel_l2": rel(logits, ref_logits)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs=2, default=[800, 1333])
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--seeds", type=int, nargs="*", default=[0])
    ap.add_argument("--modes", nargs="*", default=["bf16", "bf16+res32", "bf16bb", "bf16tr"])
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    h0, w0 = args.size
    rows = []
    for trained_like in (False, True):
        for seed in args.seeds:
            w = do.make_weights(seed, trained_like=trained_like)
            frames = do.synthetic_frames(args.frames, h0, w0, seed=1 + seed)
            t0 = time.time()
            taps32: dict = {}
            l32, b32 = do.forward(w, frames, mode="fp32", taps=taps32)
            for mode in args.modes:
                taps: dict = {}
                l, b = do.forward(w, frames, mode=mode, taps=taps)
                row = {"weights": "trained_like" if trained_like else "random_init(high gain)", "seed": seed, "frame": [h0, w0],
                       "frames": args.frames, "mode": mode, "vs": "oracle fp32", **summarise(l, b, l32, b32, h0, w0),
                       "taps_rel_l2": {k: rel(v, taps32[k]) for k, v in taps.items() if k not in ("pixel_values", "pos")}}
                rows.append(row)
                print(json.dumps({k: v for k, v in row.items() if k != "taps_rel_l2"}), flush=True)
                print("   taps:", {k: f"{v:.1e}" for k, v in row["taps_rel_l2"].items()
                                   if k in ("stem", "stage0.2", "stage1.3", "stage2.5", "stage3.2", "enc_in", "enc5", "dec0", "dec5", "dec_out")},
                      f"({time.time() - t0:.0f} s)", flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
