"""Config 5: homography + point-in-polygon + zone count over N synthetic points, every zone count / polygon kind /
output mode (bench.py --config 5 is the driver-facing line; this sweeps the variants).

Times `opd_floor_project_classify_count_f32` (the filtered kernel) with CUDA events on the launching stream around each launch,
output buffers preallocated; inputs (N x 8 B) are larger than L2 for N >= 2^25.  Prints one JSON line per configuration.
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
    ap.add_argument("--kinds", nargs="*", default=["grid", "star"])
    ap.add_argument("--modes", nargs="*", default=["idx+count", "count", "idx"])
    ap.add_argument("--probe", type=int, default=0, help="opd_set_option('probe', v): 100 + n = L2 prefetch distance n units")
    args = ap.parse_args()
    root = Path(__file__).resolve().parent.parent
    peaks = json.loads((root / "MEASURED_PEAKS.json").read_text()) if (root / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    torch.cuda.init()
    if args.probe:
        from office_person_detection_vit_b200 import _lib
        _lib.check(_lib.lib().opd_set_option(b"probe", args.probe), "probe")
    g = torch.Generator(device="cuda").manual_seed(3)
    pts = torch.empty((args.n, 2), dtype=torch.float32, device="cuda")
    pts[:, 0].uniform_(0, 1280, generator=g)
    pts[:, 1].uniform_(0, 720, generator=g)
    idx = torch.empty(args.n, dtype=torch.int32, device="cuda")
    tr = HomographyTransformer(H_CONFIG, FloorMapConfig())
    for Z in args.zones:
        for kind in args.kinds:
            zones = grid_zones(Z) if kind == "grid" else star_zones(Z, seed=4)
            zc = ZoneClassifier(zones, allow_overlap=False)
            for mode in args.modes:
                hist = torch.zeros((1, Z + 1), dtype=torch.int32, device="cuda")
                if mode == "idx+count":
                    fn = lambda: zc.count(pts, transformer=tr, out=hist, index_out=idx)  # noqa: E731
                elif mode == "count":
                    fn = lambda: zc.count(pts, transformer=tr, out=hist)  # noqa: E731
                else:
                    fn = lambda: zc._launch(pts, transformer=tr, idx_out=idx)  # noqa: E731
                for _ in range(3):
                    fn()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.iters)]
                torch.cuda.synchronize()
                for i in range(args.iters):
                    ev[2 * i].record()
                    fn()
                    ev[2 * i + 1].record()
                torch.cuda.synchronize()
                ms = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.iters))[args.iters // 2]
                bpp = {"idx+count": 12, "count": 8, "idx": 12}[mode]
                gbs = args.n * bpp / ms / 1e6
                slow = int((idx < -1).sum()) if mode != "count" else None     # placeholders must all have been overwritten
                print(json.dumps({"bench": "floor", "zones": Z, "kind": kind, "mode": mode, "n": args.n,
                                  "ms": round(ms, 4), "points_per_s": args.n / ms * 1e3, "GBps": round(gbs, 1),
                                  "frac_hbm": round(gbs / peaks["hbm_gbs"], 3), "table": zc.table.info(), "placeholders_left": slow, "probe": args.probe}))


if __name__ == "__main__":
    main()
