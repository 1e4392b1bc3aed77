"""A/B: the stage-1 identity block as the engine runs it (1x1 reduce -> fused tail, ping-pong buffers) timed per kernel,
next to the same tail launched back to back on fixed buffers (benchmarks/bneck_micro.py).  Answers whether a kernel's
in-pipeline time differs from its stand-alone time (L2 state left by the producer, sustained clocks)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200.detection import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--stage", type=int, default=0)
ap.add_argument("--set", action="append", default=[], help="library option k=v (opd_set_option)")
args = ap.parse_args()
from office_person_detection_vit_b200 import _lib  # noqa: E402
for kv in args.set:
    k, v = kv.split("=")
    _lib.check(_lib.lib().opd_set_option(k.encode(), int(v)), "opd_set_option")
B = args.batch
H, W, mid, width = ((200, 334, 64, 256), (100, 167, 128, 512))[args.stage]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, H, W, width, generator=g, device="cuda").relu().to(torch.bfloat16)
w1 = (torch.randn(mid, width, generator=g, device="cuda") / width ** 0.5).to(torch.bfloat16)
w2 = (torch.randn(mid, 3, 3, mid, generator=g, device="cuda") / (3 * mid ** 0.5)).to(torch.bfloat16)
w3 = (torch.randn(width, mid, generator=g, device="cuda") / mid ** 0.5).to(torch.bfloat16)
b1, b2, b3 = torch.zeros(mid, device="cuda"), torch.zeros(mid, device="cuda"), torch.zeros(width, device="cuda")
M = B * H * W


def block(x):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    m = ops.gemm(x.view(M, width), w1, b1, ops.EPI_BIAS_RELU).view(B, H, W, mid)
    e[1].record()
    y = ops.bottleneck_tail(m, w2, b2, w3, b3, x)
    e[2].record()
    return y, e


for _ in range(3):
    x, _e = block(x)
evs = []
for _ in range(args.iters):
    x, e = block(x)
    evs.append(e)
torch.cuda.synchronize()
t1 = sorted(e[0].elapsed_time(e[1]) for e in evs)
t2 = sorted(e[1].elapsed_time(e[2]) for e in evs)
print(f"stage{args.stage} sequence B={B}: 1x1a median {t1[len(t1) // 2]:.3f} ms (min {t1[0]:.3f}), tail median {t2[len(t2) // 2]:.3f} ms "
      f"(min {t2[0]:.3f}, max {t2[-1]:.3f})")

# the tail alone, fixed buffers, many iterations (sustained)
m = ops.gemm(x.view(M, width), w1, b1, ops.EPI_BIAS_RELU).view(B, H, W, mid)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters * 3 + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(args.iters * 3):
    y = ops.bottleneck_tail(m, w2, b2, w3, b3, x)
    ev[i + 1].record()
torch.cuda.synchronize()
t = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.iters * 3)]
print(f"  tail alone x{len(t)}: first 5 {[round(v, 3) for v in t[:5]]}, last 5 {[round(v, 3) for v in t[-5:]]}, median {sorted(t)[len(t) // 2]:.3f}")
