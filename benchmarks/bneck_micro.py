"""Microbenchmark of the fused bottleneck tail (stage-1 and stage-2 shapes of the 800x1333 forward) for ncu / A-B runs."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200 import _lib  # noqa: E402
from office_person_detection_vit_b200.detection import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--halo", type=int, default=1)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--stage", type=int, default=0)
args = ap.parse_args()
_lib.lib().opd_set_option(b"bneck_halo", args.halo)
B = args.batch
H, W, mid, width = ((200, 334, 64, 256), (100, 167, 128, 512))[args.stage]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, H, W, mid, generator=g, device="cuda").to(torch.bfloat16)
w2 = (torch.randn(mid, 3, 3, mid, generator=g, device="cuda") / (3 * mid ** 0.5)).to(torch.bfloat16)
w3 = (torch.randn(width, mid, generator=g, device="cuda") / mid ** 0.5).to(torch.bfloat16)
b2, b3 = torch.randn(mid, device="cuda"), torch.randn(width, device="cuda")
res = torch.randn(B, H, W, width, generator=g, device="cuda").to(torch.bfloat16)
for _ in range(2):
    y = ops.bottleneck_tail(x, w2, b2, w3, b3, res)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(args.iters):
    y = ops.bottleneck_tail(x, w2, b2, w3, b3, res)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.iters))[args.iters // 2]
M = B * H * W
gb = (2 * M * mid + 4 * M * width) / 1e9
tf = 2 * M * mid * (9 * mid + width) / 1e12
print(f"stage{args.stage} halo={args.halo} B={B}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  {tf / ms * 1e3:.0f} TF/s  ({ms / 64 * B * 64 / B:.3f} ms per 64-frame layer x{64 // B})")
