"""Instruction-pipe throughput on one SM (csrc/probe/pipe_probe.cu, libopd_probe.so): thread-level operations per clock per SM for
MUFU.EX2, cvt.rn.bf16x2.f32, fma.rn.f32x2 and the attention kernel's softmax mix, at 8 / 16 / 32 warps per SM."""
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200 import _lib  # noqa: E402

_lib.lib()
_probe = C.CDLL(str(_lib.LIB_PATH.with_name("libopd_probe.so")))
_probe.opd_debug_pipe_probe.restype = C.c_int
_probe.opd_debug_pipe_probe.argtypes = [C.c_int32] * 4 + [C.c_void_p, C.c_void_p, C.c_void_p]

torch.cuda.init()
sms = torch.cuda.get_device_properties(0).multi_processor_count
out = torch.zeros(4, device="cuda")
iters = 20000
for mode, name, per_iter in ((0, "ex2.approx.ftz.f32", 8), (1, "cvt.rn.bf16x2.f32", 8), (2, "fma.rn.f32x2 (pairs)", 4), (3, "softmax mix (element pairs)", 4)):
    for ctas_per_sm, threads in ((1, 256), (2, 256), (4, 256), (8, 256)):
        grid = sms * ctas_per_sm
        cyc = torch.zeros(grid, dtype=torch.int64, device="cuda")
        for _ in range(2):
            _lib.check(_probe.opd_debug_pipe_probe(mode, iters, grid, threads, out.data_ptr(), cyc.data_ptr(), None), "probe")
        torch.cuda.synchronize()
        c = float(cyc.double().mean())
        per_clk_sm = iters * per_iter * threads * ctas_per_sm / c
        print(json.dumps({"op": name, "warps_per_sm": ctas_per_sm * threads // 32, "cycles": c,
                          "thread_ops_per_clk_per_sm": round(per_clk_sm, 2)}), flush=True)
