"""Per-launch times of one 800x1333 forward (opd_detr_profile) with library options set, for A/B probes:
python benchmarks/step_times.py --batch 64 --set probe=1 --filter stem"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200 import _lib  # noqa: E402
from office_person_detection_vit_b200.detection import ViTDetector  # noqa: E402
from oracle import detr_oracle as do  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--set", action="append", default=[])
ap.add_argument("--filter", default="")
ap.add_argument("--size", default="800x1333")
args = ap.parse_args()
h, w = (int(v) for v in args.size.split("x"))
det = ViTDetector(state_dict=do.make_weights(0))
det.load_model()
frames = torch.randint(0, 256, (args.batch, h, w, 3), dtype=torch.uint8, device="cuda")
for variant in [[]] + [[s] for s in args.set]:
    for name in ("probe",):
        _lib.lib().opd_set_option(name.encode(), 0)
    for s in variant:
        k, v = s.split("=")
        _lib.check(_lib.lib().opd_set_option(k.encode(), int(v)), "opd_set_option")
    det.model.set_debug(False)   # drops the cached launch plan: plan-time options take effect
    for _ in range(2):
        det.model.forward(frames)
    torch.cuda.synchronize()
    profs = [det.model.profile() for _ in range(3)]
    rows = profs[0]
    tot = 0.0
    out = []
    for i, r in enumerate(rows):
        ms = sorted(p[i]["ms"] for p in profs)[1]
        tot += ms
        if args.filter in r["name"]:
            out.append(f"{r['name']}={ms:.3f}")
    print(f"{variant or 'default'}: total {tot:.3f} ms; " + " ".join(out))
