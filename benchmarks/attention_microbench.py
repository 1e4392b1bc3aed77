"""attention_tc_kernel variants on the shapes of the batch-64 forward (encoder self-attention, decoder cross- and self-attention).

    python benchmarks/attention_microbench.py [--iters 10] [--kv 64 65 128]

kv = opd_set_option("attention_kv"): 96 (default) / 64 / 128 keys per tile, 65 = 64 keys with the S tile held in registers (early S issue).
CUDA events on the launching stream around every launch, median; outputs compared bit for bit with kv = 64."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200 import _lib  # noqa: E402
from office_person_detection_vit_b200.detection import ops  # noqa: E402

SHAPES = (("enc.self", 64, 1050, 1050), ("dec.cross", 64, 100, 1050), ("dec.self", 64, 100, 100), ("enc.self@B8", 8, 1050, 1050))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--kv", type=int, nargs="*", default=[96, 64, 65, 128])
    args = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, B, Lq, Lk in SHAPES:
        q = (torch.randn(B, Lq, 256, generator=g, device="cuda") * 1.5).to(torch.bfloat16)
        k = (torch.randn(B, Lk, 256, generator=g, device="cuda") * 1.5).to(torch.bfloat16)
        v = torch.randn(B, Lk, 256, generator=g, device="cuda").to(torch.bfloat16)
        base = None
        for kv in args.kv:
            _lib.check(_lib.lib().opd_set_option(b"attention_kv", kv), "attention_kv")
            for _ in range(2):
                o = ops.attention(q, k, v, 8)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.iters)]
            torch.cuda.synchronize()
            for i in range(args.iters):
                ev[2 * i].record()
                o = ops.attention(q, k, v, 8)
                ev[2 * i + 1].record()
            torch.cuda.synchronize()
            ms = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.iters))[args.iters // 2]
            if base is None:
                base = o.clone()
            print(json.dumps({"shape": name, "B": B, "Lq": Lq, "Lk": Lk, "kv": kv, "us": round(ms * 1e3, 1),
                              "tflops": round(4 * B * 8 * Lq * Lk * 32 / (ms * 1e-3) / 1e12, 1),
                              "bit_identical_to_first": bool(torch.equal(o, base)),
                              "max_abs_diff_to_first": float((o.float() - base.float()).abs().max())}), flush=True)
        _lib.lib().opd_set_option(b"attention_kv", 96)


if __name__ == "__main__":
    main()
