import ctypes as C, torch, sys
sys.path.insert(0, "/root/repo")
from office_person_detection_vit_b200 import _lib
P = C.c_void_p
_lib.lib()   # the probe library (csrc/probe/, libopd_probe.so) links against libopd_b200.so
_probe = C.CDLL(str(_lib.LIB_PATH.with_name("libopd_probe.so")))
_probe.opd_halo_conv3x3_test.restype = C.c_int
_probe.opd_halo_conv3x3_test.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, P, P, P, C.c_int32, P]
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(0)
for (B, H, W) in [(1, 16, 8), (2, 37, 29), (1, 200, 334)]:
    x = torch.randn(B, H, W, 64, generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn(64, 3, 3, 64, generator=g, device="cuda") / 24).to(torch.bfloat16)
    bias = torch.randn(64, device="cuda")
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1).relu().permute(0, 2, 3, 1)
    for mode in (0, 1):
        y = torch.zeros(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
        rc = _probe.opd_halo_conv3x3_test(x.data_ptr(), B, H, W, w.data_ptr(), bias.data_ptr(), y.data_ptr(), mode, _lib.stream_ptr())
        assert rc == 0, _lib.lib().opd_last_error()
        torch.cuda.synchronize()
        err = (y.float() - ref).abs()
        print((B, H, W), "mode", mode, "max err", float(err.max()), "frac bad", float((err > 0.05 + 0.02 * ref.abs()).float().mean()))
