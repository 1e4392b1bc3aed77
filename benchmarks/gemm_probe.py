"""In-kernel cycle counters of tc_gemm_kernel (build with OPD_EXTRA_NVCC_FLAGS=-DOPD_GEMM_PROBE): where CTA 0's MMA thread and
epilogue spend their time for the layer shapes of the batch-64 forward."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from office_person_detection_vit_b200.detection import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
for name, M, N, K, epi in (("enc.fc1", 67200, 2048, 256, 1), ("enc.fc2 (no LN)", 67200, 256, 2048, 0), ("stage3 1x1 expand", 268800, 1024, 256, 2),
                           ("stage3 1x1 reduce", 268800, 256, 1024, 1), ("enc.qk", 67200, 512, 256, 0),
                           ("enc.o+ln", 67200, 256, 256, 3), ("enc.fc2+ln", 67200, 256, 2048, 3), ("dec.so+ln", 6400, 256, 256, 3),
                           ("dec.fc2+ln", 6400, 256, 2048, 3), ("dec.fc1", 6400, 2048, 256, 1), ("dec.cq", 6400, 256, 256, 0)):
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, generator=g, device="cuda").to(torch.bfloat16) if epi >= 2 else None
    kw = dict(gamma=torch.randn(N, device="cuda"), beta=torch.randn(N, device="cuda")) if epi == 3 else {}
    for _ in range(2):
        ops.gemm(a, w, bias, epilogue=epi, residual=res, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print(f"--- {name}: M={M} N={N} K={K}", flush=True)
    e0.record()
    ops.gemm(a, w, bias, epilogue=epi, residual=res, **kw)
    e1.record()
    torch.cuda.synchronize()
    print(f"    {e0.elapsed_time(e1) * 1e3:.1f} us, {2 * M * N * K / e0.elapsed_time(e1) / 1e9:.0f} TF/s", flush=True)
