"""tc_gemm_kernel on the GEMM shapes of the batch-64 forward under plan-selection options (A/B runs for finish_plan's rules).

    python benchmarks/gemm_shapes.py [--iters 10] [--opts default gemm_outbufs=0 gemm_res_wide=1 ...]

Each shape runs `iters` times back to back (operands + outputs exceed L2 for the large shapes), CUDA events on the launching
stream around every launch, median reported with both roofline fractions (MEASURED_PEAKS.json).  The outputs of every option set
are compared bit for bit with the default's."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200 import _lib  # noqa: E402
from office_person_detection_vit_b200.detection import ops  # noqa: E402

SHAPES = (("stage3.conv1x1b", 67200, 2048, 512, 2), ("stage2.conv1x1b", 268800, 1024, 256, 2), ("stage1.conv1x1b-like", 268800, 512, 128, 2),
          ("dec.cross_kv", 67200, 1536, 256, 0), ("enc.fc1", 67200, 2048, 256, 1), ("enc.fc2+ln", 67200, 256, 2048, 3),
          ("input_proj", 67200, 256, 2048, 0), ("enc.qk", 67200, 512, 256, 0), ("enc.o+ln", 67200, 256, 256, 3),
          ("stage3.conv1x1a", 67200, 512, 2048, 1), ("stage2.conv1x1a", 268800, 256, 1024, 1),
          ("stage2.conv1x1b@B8", 33600, 1024, 256, 2), ("stage3.conv1x1b@B8", 8400, 2048, 512, 2),
          ("dec.fc2+ln", 6400, 256, 2048, 3), ("dec.fc1", 6400, 2048, 256, 1), ("stage4.conv1x1a", 16800, 512, 2048, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--opts", nargs="*", default=["default", "gemm_outbufs=0", "gemm_res_wide=0", "gemm_res_wide=2"])
    ap.add_argument("--shapes", nargs="*", default=None)
    args = ap.parse_args()
    root = Path(__file__).resolve().parent.parent
    peaks = json.loads((root / "MEASURED_PEAKS.json").read_text()) if (root / "MEASURED_PEAKS.json").exists() else {}
    hbm = peaks.get("hbm_gbs", 6650.0) * 1e9
    tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1380.0)) * 1e12
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, M, N, K, epi in SHAPES:
        if args.shapes and not any(s in name for s in args.shapes):
            continue
        a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, generator=g, device="cuda").to(torch.bfloat16) if epi >= 2 else None
        kw = dict(gamma=torch.randn(N, device="cuda"), beta=torch.randn(N, device="cuda")) if epi == 3 else {}
        nbytes = 2 * (M * K + N * K + M * N * (2 if epi >= 2 else 1))
        base = None
        for opt in args.opts:
            if opt != "default":
                k, v = opt.split("=")
                _lib.check(_lib.lib().opd_set_option(k.encode(), int(v)), opt)
            for _ in range(2):
                d = ops.gemm(a, w, bias, epilogue=epi, residual=res, **kw)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.iters)]
            torch.cuda.synchronize()
            for i in range(args.iters):
                ev[2 * i].record()
                d = ops.gemm(a, w, bias, epilogue=epi, residual=res, **kw)
                ev[2 * i + 1].record()
            torch.cuda.synchronize()
            ms = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.iters))[args.iters // 2]
            same = None
            if base is None:
                base = d.clone()
            else:
                same = bool(torch.equal(d, base))
            print(json.dumps({"shape": name, "M": M, "N": N, "K": K, "epi": epi, "opt": opt, "us": round(ms * 1e3, 1),
                              "frac_tensor": round(2 * M * N * K / (ms * 1e-3) / tf, 3), "frac_hbm": round(nbytes / (ms * 1e-3) / hbm, 3),
                              "bit_identical_to_default": same}), flush=True)
            if opt != "default":
                _lib.check(_lib.lib().opd_set_option(k.encode(), {"gemm_outbufs": 1, "gemm_res_wide": 1}.get(k, 0)), opt)
            del d
        del a, w, res, base


if __name__ == "__main__":
    main()
