"""Timeline of CTA 0 of tc_bneck_kernel<128> (globaltimer stamps at the pipeline hand-offs).
Needs a measurement build: OPD_EXTRA_NVCC_FLAGS="-DOPD_BNECK_PROBE" python -m office_person_detection_vit_b200.build -f"""
import ctypes as C, sys, collections
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from office_person_detection_vit_b200 import _lib
from office_person_detection_vit_b200.detection import ops
_lib.register("opd_debug_set_bneck_trace", C.c_int, [C.c_void_p])
_lib.lib().opd_set_option(b"bneck_halo", 1)
B, H, W, mid, width = 16, 100, 167, 128, 512
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, H, W, mid, generator=g, device="cuda").to(torch.bfloat16)
w2 = (torch.randn(mid, 3, 3, mid, generator=g, device="cuda") / (3 * mid ** 0.5)).to(torch.bfloat16)
w3 = (torch.randn(width, mid, generator=g, device="cuda") / mid ** 0.5).to(torch.bfloat16)
b2, b3 = torch.randn(mid, device="cuda"), torch.randn(width, device="cuda")
res = torch.randn(B, H, W, width, generator=g, device="cuda").to(torch.bfloat16)
for _ in range(2):
    ops.bottleneck_tail(x, w2, b2, w3, b3, res)
tr = torch.zeros(4096, dtype=torch.int64, device="cuda")
_lib.lib().opd_debug_set_bneck_trace(tr.data_ptr())
ops.bottleneck_tail(x, w2, b2, w3, b3, res)
torch.cuda.synchronize()
_lib.lib().opd_debug_set_bneck_trace(None)
t = tr.cpu().numpy()
n = int(t[0]); recs = sorted(((int(v) & ((1 << 48) - 1), int(v) >> 48) for v in t[1:1 + min(n, 4000)]))
t0 = recs[0][0]
names = {1: "P:g1 loads issued", 2: "P:g2 loads issued", 11: "M:g1 issued", 12: "M:g2 issued", 13: "M:a2_ready seen",
         20: "E1:begin", 21: "E1:acc1_full", 22: "E1:a2_free", 23: "E1:done", 30: "E2:begin", 31: "E2:acc2_full", 32: "E2:res_full", 33: "E2:done"}
print("events", n)
for ts, ev in recs[60:200]:
    print(f"{(ts - t0) / 1000:9.2f} us  {names.get(ev, ev)}")
# average durations between consecutive epilogue events
seq = [(ts, ev) for ts, ev in recs if ev >= 20]
acc = collections.defaultdict(list)
for (ta, ea), (tb, eb) in zip(seq, seq[1:]):
    acc[(ea, eb)].append(tb - ta)
for k, v in sorted(acc.items()):
    print(names[k[0]], "->", names[k[1]], f"n={len(v)} mean {sum(v) / len(v) / 1000:.3f} us")
