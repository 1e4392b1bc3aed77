"""Cycles per tcgen05.mma (128 x N x 16) by shape, swizzle, accumulator rotation and A-operand view (csrc/probe/mma_probe.cu,
built into libopd_probe.so - not part of the product library)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200 import _lib  # noqa: E402
import ctypes as C  # noqa: E402

_lib.lib()   # the probe library links against libopd_b200.so
_probe = C.CDLL(str(_lib.LIB_PATH.with_name("libopd_probe.so")))
_probe.opd_debug_mma_probe.restype = C.c_int
_probe.opd_debug_mma_probe.argtypes = [C.c_int32] * 10 + [C.c_void_p, C.c_void_p]

out = torch.zeros(296, dtype=torch.int64, device="cuda")
iters, grid = 2048, 148


def run(N, sw32, n_acc, walk, sbo=0, step=0, ld_iters=0, both=False, commit_every=0):
    for _ in range(2):
        _lib.check(_probe.opd_debug_mma_probe(N, sw32, n_acc, iters, walk, grid, sbo, step, ld_iters, commit_every, out.data_ptr(), None), "probe")
    torch.cuda.synchronize()
    mma = out[:grid].float().median().item() / iters
    if both:
        return mma, out[grid:2 * grid].float().median().item() / max(ld_iters, 1)
    return mma


print("N  swizzle  n_acc  cycles/MMA   (math floor N/2; operand floor (4096 + 32 N) / 128)")
for sw32 in (0, 1):
    for N in (32, 64, 128, 256):
        for n_acc in (1, 2):
            print(f"{N:3d}  {'32B ' if sw32 else '128B'}  {n_acc}  {run(N, sw32, n_acc, 1):7.1f}   {N / 2:.0f}  {(4096 + 32 * N) / 128:.0f}")
print("shifted views (A = halo patch): ")
print(f"stem   N=64  32B-swizzle  sbo 352  step 32:   {run(64, 1, 2, 0, 352, 32):7.1f}")
print(f"stem   N=64  32B-swizzle  sbo 384  step 32:   {run(64, 1, 2, 0, 384, 32):7.1f}")
print(f"stem   N=64  32B-swizzle  sbo 256  step 32:   {run(64, 1, 2, 0, 256, 32):7.1f}")
print(f"halo   N=64  128B-swizzle sbo 2304 step 128:  {run(64, 0, 2, 0, 2304, 128):7.1f}")
print(f"halo   N=64  128B-swizzle sbo 2304 step 32:   {run(64, 0, 2, 0, 2304, 32):7.1f}")
print(f"halo   N=128 128B-swizzle sbo 2304 step 128:  {run(128, 0, 2, 0, 2304, 128):7.1f}")
print(f"plain  N=64  128B-swizzle sbo 1024 step 32:   {run(64, 0, 2, 0, 1024, 32):7.1f}")
print("TMEM reads against the MMAs (4 warps x tcgen05.ld 32x32b.x32 + wait = 16 KB per round):")
for N, n_ld in ((64, 2048 * 48 // 150), (256, 2048 * 128 // 150), (64, 16), (256, 16)):
    mma, ld = run(N, 0, 2, 1, ld_iters=n_ld, both=True)
    print(f"  N={N:3d}: {mma:6.1f} cycles / MMA with {n_ld} loads per warp in flight, {ld:6.1f} cycles per load round")
print("two tcgen05.commit every k MMAs (N = 64):")
for k in (8, 16, 64, 256):
    print(f"  every {k:3d} MMAs: {run(64, 0, 2, 1, commit_every=k):6.1f} cycles / MMA")
