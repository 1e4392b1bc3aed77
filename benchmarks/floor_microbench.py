"""Config 5: homography + point-in-polygon + zone count over N synthetic points.

Times `opd_floor_project_classify_count_f32` (the filtered kernel) with CUDA events on the launching
stream; inputs (N x 8 B) are larger than L2 for N >= 2^25.  Prints one JSON line per configuration.
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer  # noqa: E402
from office_person_detection_vit_b200.zone import ZoneClassifier  # noqa: E402
from office_person_detection_vit_b200.scene import H_CONFIG, grid_zones, star_zones  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--zones", type=int, nargs="*", default=[4, 16, 64])
    args = ap.parse_args()
    peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) \
        if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    torch.cuda.init()
    g = torch.Generator(device="cuda").manual_seed(3)
    pts = torch.empty((args.n, 2), dtype=torch.float32, device="cuda")
    pts[:, 0].uniform_(0, 1280, generator=g)
    pts[:, 1].uniform_(0, 720, generator=g)
    tr = HomographyTransformer(H_CONFIG, FloorMapConfig())
    for Z in args.zones:
        for kind in ("grid", "star"):
            zones = grid_zones(Z) if kind == "grid" else star_zones(Z, seed=4)
            zc = ZoneClassifier(zones, allow_overlap=False)
            for mode in ("idx+count", "count"):
                hist = torch.zeros((1, Z + 1), dtype=torch.int32, device="cuda")
                fn = (lambda: zc.count(pts, transformer=tr, out=hist, return_index=True)) if mode == "idx+count" \
                    else (lambda: zc.count(pts, transformer=tr, out=hist))
                for _ in range(3):
                    fn()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters + 1)]
                torch.cuda.synchronize()
                ev[0].record()
                for i in range(args.iters):
                    fn()
                    ev[i + 1].record()
                torch.cuda.synchronize()
                ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.iters))[args.iters // 2]
                bpp = 12 if mode == "idx+count" else 8
                gbs = args.n * bpp / ms / 1e6
                print(json.dumps({"bench": "floor", "zones": Z, "kind": kind, "mode": mode, "n": args.n,
                                  "ms": round(ms, 4), "points_per_s": args.n / ms * 1e3, "GBps": round(gbs, 1),
                                  "frac_hbm": round(gbs / peaks["hbm_gbs"], 3), "table": zc.table.info()}))


if __name__ == "__main__":
    main()
