"""Summary of an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of bench.py:
one period of the step (preprocess kernel to preprocess kernel), per-kernel time share and DRAM traffic, and the JSON that
bench.py reads for roofline.traffic.   python benchmarks/summarize_launches.py launches.csv plain.log out_prefix"""
import collections
import csv
import json
import re
import sys

csv_path, plain_log, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
line = [l for l in open(plain_log) if l.startswith("{")][-1]
d = json.loads(line)
P = d["gpu_launches"] // d["steps"]
rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, mi, vi, ii, ui, gi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit", "Grid Size"))
L = collections.OrderedDict()
for r in rows:
    dd = L.setdefault(r[ii], {"name": r[ki], "grid": r[gi]})
    dd[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
allL = list(L.values())
# the step is periodic in the launch list: P library launches + torch's fill kernels (hist.zero_, output buffers).  The period
# is the smallest shift >= P under which the captured names repeat; any window of that length is one whole step (rotated).
names = [x["name"] for x in allL]
# (a plan's first call launches its frame-independent prologue once: the periodic part may start a few launches in)
s0, Pn = next((s, k) for s in range(0, 64) for k in range(P, (len(names) - s) // 2 + 1)
              if all(names[i] == names[i + k] for i in range(s, len(names) - k)))
step = allL[s0:s0 + Pn]

def short(n):
    n = re.sub(r"\(.*", "", n)
    return n.replace("void ", "").replace("opd::<unnamed>::", "").replace("opd::", "").replace("<unnamed>::", "")


def val(x, k):
    v, u = x[k]
    if k == "gpu__time_duration.sum":
        return v / 1e3 if u == "ns" else v
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


agg = collections.OrderedDict()
for x in step:
    a = agg.setdefault(short(x["name"]), [0, 0.0, 0.0, 0.0, x["grid"]])
    a[0] += 1
    a[1] += val(x, "gpu__time_duration.sum")
    a[2] += val(x, "dram__bytes_read.sum")
    a[3] += val(x, "dram__bytes_write.sum")
tot = sum(a[1] for a in agg.values())
out = [f"{'kernel':64s}   n      us  share  dramR GB dramW GB  grid"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k[:64]:64s} {a[0]:3d} {a[1]:8.1f} {100 * a[1] / tot:5.1f}% {a[2] / 1e9:7.2f} {a[3] / 1e9:7.2f}  {a[4]}")
fam = [a for k, a in agg.items() if any(s in k for s in ("tc_gemm_kernel", "tc_mlp_kernel", "tc_bneck", "stem_kernel"))]
out.append(f"tensor-core family: {sum(a[0] for a in fam)} launches, {sum(a[1] for a in fam):.0f} us = {100 * sum(a[1] for a in fam) / tot:.1f}% "
           f"of the step under ncu, DRAM {sum(a[2] + a[3] for a in fam) / 1e9:.2f} GB")
out.append(f"whole step: {Pn} launches ({P} of this library + {Pn - P} torch fills), {tot:.0f} us under ncu (serialised, cold caches), DRAM {sum(a[2] + a[3] for a in agg.values()) / 1e9:.2f} GB; "
           f"the same command without ncu: {d['ms_per_step']} ms per step")
print("\n".join(out))
open(prefix + "_summary.txt", "w").write("\n".join(out) + "\n")
json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one period of "
                     "`python bench.py --steps 2 --warmup 3 --no-graph` at batch 64 (" + prefix + ".csv)",
           "launches_per_step": Pn, "family_launches": sum(a[0] for a in fam), "family_dram_bytes_per_step": sum(a[2] + a[3] for a in fam),
           "family_share_of_step_under_ncu": round(sum(a[1] for a in fam) / tot, 4), "step_dram_bytes": sum(a[2] + a[3] for a in agg.values())},
          open(prefix.replace("launches_final_b64", "traffic") + ".json", "w"), indent=1)
