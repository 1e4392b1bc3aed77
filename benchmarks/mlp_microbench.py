"""Fused feed-forward block (tc_mlp.cu) against the two GEMM launches it replaces, at the encoder (M = 67 200) and decoder (M = 6 400)
row counts of the batch-64 forward.  CUDA events on the launching stream around every call, median of `iters`."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200 import _lib  # noqa: E402
from office_person_detection_vit_b200.detection import ops  # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)]
    torch.cuda.synchronize()
    for i in range(iters):
        ev[2 * i].record()
        fn()
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(iters))[iters // 2] * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g, device="cuda") * scale).to(torch.bfloat16)  # noqa: E731
    for name, M, pos_rows in (("encoder", 67200, 1050), ("decoder", 6400, 100), ("encoder@B8", 8400, 1050)):
        x, w1, w2 = rnd(M, 256), rnd(2048, 256, scale=1 / 16), rnd(256, 2048, scale=1 / 45)
        b1, b2 = torch.randn(2048, device="cuda") * 0.1, torch.randn(256, device="cuda") * 0.1
        gamma, beta, pos = torch.rand(256, device="cuda") + 0.5, torch.randn(256, device="cuda") * 0.1, torch.randn(pos_rows, 256, device="cuda")

        def two():
            h = ops.gemm(x, w1, b1, epilogue=1)
            return ops.gemm(h, w2, b2, epilogue=3, residual=x, gamma=gamma, beta=beta, pos=pos)

        t2 = timed(two, args.iters)
        a = two()
        for pair in (1, 0):
            _lib.check(_lib.lib().opd_set_option(b"mlp_pair", pair), "mlp_pair")
            t1 = timed(lambda: ops.mlp_ln(x, w1, b1, w2, b2, gamma, beta, pos=pos), args.iters)
            b = ops.mlp_ln(x, w1, b1, w2, b2, gamma, beta, pos=pos)
            print(json.dumps({"shape": name, "M": M, "two_gemms_us": round(t2, 1), "fused_us": round(t1, 1), "pair": pair,
                              "fused_tflops": round(4 * M * 256 * 2048 / t1 / 1e6, 1),
                              "bit_identical": bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]))}), flush=True)
        _lib.lib().opd_set_option(b"mlp_pair", 1)


if __name__ == "__main__":
    main()
