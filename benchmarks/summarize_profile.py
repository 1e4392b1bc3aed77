"""Summarise a per-launch timing table written by `bench.py --profile-out` (opd_detr_profile)."""
import collections
import json
import sys

p = json.load(open(sys.argv[1]))
tot = sum(s["ms"] for s in p)
print(f"total {tot:.3f} ms over {len(p)} launches")
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for s in p:
    n = s["name"]
    if n.startswith("stage"):
        key = n.split(".")[0] + ":" + n.split(".")[-1]
    elif n.startswith("enc") and "." in n:
        key = "enc:" + n.split(".", 1)[1]
    elif n.startswith("dec") and n[3:4].isdigit():
        key = "dec:" + s["kind"]
    else:
        key = n
    a = agg[key]
    a[0] += s["ms"]; a[1] += s["flops"]; a[2] += s["bytes"]; a[3] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:26s} n={v[3]:3d} {v[0]:7.3f} ms {100 * v[0] / tot:5.1f}%  {v[1] / v[0] / 1e9:7.1f} TF/s {v[2] / v[0] / 1e6:7.0f} GB/s")
