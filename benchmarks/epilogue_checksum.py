"""SHA-256 of the outputs of the tensor-core building blocks and of one whole detector forward on fixed seeded inputs: run
before and after a kernel change that must not change a bit (epilogue instruction selection, pipelining, scheduling)."""
import hashlib
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200.detection import ViTDetector, ops  # noqa: E402
from office_person_detection_vit_b200.detection.synthetic import random_init_state_dict, synthetic_frames  # noqa: E402


def rnd(*shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def digest(t):
    return hashlib.sha256(t.contiguous().view(torch.uint8).cpu().numpy().tobytes()).hexdigest()[:16]


out = {}
g = torch.Generator(device="cuda").manual_seed(99)
for (B, H, W, mid, width, stride) in [(2, 50, 83, 64, 256, 1), (3, 40, 67, 128, 512, 1), (2, 51, 84, 128, 512, 2), (2, 64, 64, 128, 512, 1)]:
    x = rnd(B, H, W, mid, seed=20)
    x[0, 0, 0, :8] = float("nan")      # NaN / -0.0 handling must not change either
    w2, w3 = rnd(mid, 3, 3, mid, seed=21, scale=(9 * mid) ** -0.5), rnd(width, mid, seed=22, scale=mid ** -0.5)
    b2, b3 = torch.randn(mid, device="cuda", generator=g) * 0.3, torch.randn(width, device="cuda", generator=g) * 0.3
    P, Q = (H - 1) // stride + 1, (W - 1) // stride + 1
    res = rnd(B, P, Q, width, seed=23)
    out[f"tail {B}x{H}x{W} mid{mid} s{stride}"] = digest(ops.bottleneck_tail(x, w2, b2, w3, b3, res, stride=stride))
for (M, N, K) in [(300, 64, 128), (2100, 256, 256), (1000, 2048, 256), (128 * 150 + 7, 128, 576), (6400, 512, 1024), (128 * 300, 1024, 256)]:
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda", generator=g)
    res = rnd(M, N, seed=3)
    for epi in (0, 1, 2):
        out[f"gemm {M}x{N}x{K} epi{epi}"] = digest(ops.gemm(a, w, bias, epilogue=epi, residual=res if epi == 2 else None))
    if N == 256:
        gamma, beta = torch.randn(N, device="cuda", generator=g), torch.randn(N, device="cuda", generator=g)
        out[f"gemm {M}x{N}x{K} ln"] = digest(ops.gemm(a, w, bias, epilogue=3, residual=res, gamma=gamma, beta=beta))
x = rnd(2, 40, 52, 256, seed=5)
w = rnd(256, 3, 3, 256, seed=6, scale=(9 * 256) ** -0.5)
out["conv3x3"] = digest(ops.conv2d_nhwc(x, w, torch.randn(256, device="cuda", generator=g), stride=1, pad=1, epilogue=1))
det = ViTDetector(confidence_threshold=0.5, state_dict=random_init_state_dict(0), device="cuda:0", batch_size=4)
det.load_model()
frames = torch.from_numpy(synthetic_frames(4, 800, 1333, seed=7)).cuda()
o = det.detect_tensors(frames, bgr=True)
torch.cuda.synchronize()
for k in sorted(o):
    if hasattr(o[k], "is_cuda"):
        out[f"detect {k}"] = digest(o[k])
print(json.dumps(out, indent=1))
