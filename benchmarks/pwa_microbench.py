"""Piecewise-affine transform throughput: N points through opd_pwa_transform_f64 (24 correspondences -> 37 triangles, the golden
set), CUDA events, median of `iters`.  Algorithmic traffic: 16 B read + 16 B written per point."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from office_person_detection_vit_b200.transform import FloorMapConfig, PiecewiseAffineTransformer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=50_000_000)
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
root = Path(__file__).resolve().parent.parent
g = np.load(root / "tests" / "golden" / "pwa_golden.npz")
peaks = json.loads((root / "MEASURED_PEAKS.json").read_text()) if (root / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
tr = PiecewiseAffineTransformer(g["src"], g["dst"], FloorMapConfig())
gen = torch.Generator(device="cuda").manual_seed(3)
pts = torch.empty((args.n, 2), dtype=torch.float64, device="cuda")
pts[:, 0].uniform_(0, 1280, generator=gen)
pts[:, 1].uniform_(0, 720, generator=gen)
for _ in range(2):
    tr.transform_points(pts)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(args.iters):
    tr.transform_points(pts)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.iters))[args.iters // 2]
gbs = args.n * 32 / ms / 1e6
print(json.dumps({"bench": "pwa", "triangles": len(tr.affine_matrices), "n": args.n, "ms": round(ms, 4), "points_per_s": args.n / ms * 1e3,
                  "GBps": round(gbs, 1), "frac_hbm": round(gbs / peaks["hbm_gbs"], 3)}))
