"""GPU numerics of the tensor-core building blocks (csrc/tc_gemm.cu, attention) through the C ABI, against a plain
PyTorch fp32 reference of the same op on the same bf16-rounded operands.  Tolerance: the outputs are bf16
(8 significant bits), the accumulation is fp32 on both sides -> |err| <= 2^-8 * |ref| + small absolute term."""

from __future__ import annotations

import pytest

pytestmark = pytest.mark.gpu
PAIR_DEFAULT = 1   # library default of opd_set_option("gemm_pair")
BNECK_PAIR_DEFAULT = 1   # library default of opd_set_option("bneck_pair")


@pytest.fixture(scope="module")
def T(built_lib):
    import torch

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch


def _close(torch, got, ref, what):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2 * ref.abs().mean().clamp_min(1e-3)
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.numel()} off, max err {float(err.max()):.4g}, " \
                          f"ref scale {float(ref.abs().mean()):.4g}, first bad {bad.nonzero()[:4].tolist()}"


def _rand(torch, *shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 64, 128), (2100, 256, 256), (1000, 2048, 256),
                                   (1000, 256, 2048), (128 * 150 + 7, 128, 576), (6400, 512, 1024)])
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_gemm_epilogues(T, M, N, K, epi):
    from office_person_detection_vit_b200.detection import ops

    torch = T
    a, w = _rand(torch, M, K, seed=1), _rand(torch, N, K, seed=2, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    res = _rand(torch, M, N, seed=3) if epi == 2 else None
    d = ops.gemm(a, w, bias, epilogue=epi, residual=res)
    ref = a.float() @ w.float().T + bias
    if epi == 2:
        ref = ref + res.float()
    if epi >= 1:
        ref = ref.relu()
    _close(torch, d, ref, f"gemm {M}x{N}x{K} epi {epi}")


@pytest.mark.parametrize("M,N,K", [(128 * 40 + 5, 1024, 256), (128 * 21, 512, 128), (128 * 300 + 77, 256, 256), (200, 1024, 64)])
def test_gemm_weight_stationary_variant_is_bit_identical(T, M, N, K):
    """The weight-stationary kernel variant (one n-block per CTA, its weights resident in shared memory; chosen for the
    short-K bottleneck outputs of the batch-64 forward) accumulates in the same order as the streaming variant: the
    outputs must be bit-identical, whatever the number of m-blocks per CTA (including CTAs without work)."""
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ops

    torch = T
    a, w = _rand(torch, M, K, seed=11), _rand(torch, N, K, seed=12, scale=K ** -0.5)
    bias, res = torch.randn(N, device="cuda"), _rand(torch, M, N, seed=13)
    try:
        _lib.check(_lib.lib().opd_set_option(b"gemm_res_wide", 0), "opd_set_option")   # the variant belongs to the 128-column tiles
        _lib.check(_lib.lib().opd_set_option(b"gemm_bres", 0), "opd_set_option")
        ref = ops.gemm(a, w, bias, epilogue=2, residual=res)
        _lib.check(_lib.lib().opd_set_option(b"gemm_bres", 2), "opd_set_option")
        got = ops.gemm(a, w, bias, epilogue=2, residual=res)
        torch.cuda.synchronize()
    finally:
        _lib.lib().opd_set_option(b"gemm_bres", 1)
        _lib.lib().opd_set_option(b"gemm_res_wide", 1)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    _close(torch, got, (a.float() @ w.float().T + bias + res.float()).relu(), f"weight-stationary gemm {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", [(128 * 301 + 77, 2048, 512), (128 * 160, 1024, 256), (128 * 9 + 1, 512, 256), (128 * 75, 256, 128), (90, 1024, 512)])
def test_gemm_residual_wide_tiles_are_bit_identical(T, M, N, K):
    """Bias + residual + ReLU layers with K >= 256 run on 256-column cta_group::2 tiles with a three-slot residual ring (default);
    the 128-column kernel they replaced accumulates in the same order: bit-identical outputs for even / odd / ragged m-block
    counts, for shapes with too few tile pairs for clusters (single-CTA 256-column kernel) and when forced on a K = 128 layer."""
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ops

    torch = T
    a, w = _rand(torch, M, K, seed=41), _rand(torch, N, K, seed=42, scale=K ** -0.5)
    bias, res = torch.randn(N, device="cuda"), _rand(torch, M, N, seed=43)
    try:
        _lib.check(_lib.lib().opd_set_option(b"gemm_res_wide", 0), "opd_set_option")
        ref = ops.gemm(a, w, bias, epilogue=2, residual=res)
        _lib.check(_lib.lib().opd_set_option(b"gemm_res_wide", 2), "opd_set_option")
        got = ops.gemm(a, w, bias, epilogue=2, residual=res)
        _lib.check(_lib.lib().opd_set_option(b"gemm_res_wide", 1), "opd_set_option")
        dflt = ops.gemm(a, w, bias, epilogue=2, residual=res)
        torch.cuda.synchronize()
    finally:
        _lib.lib().opd_set_option(b"gemm_res_wide", 1)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    assert torch.equal(dflt.view(torch.int16), ref.view(torch.int16))
    _close(torch, got, (a.float() @ w.float().T + bias + res.float()).relu(), f"wide residual gemm {M}x{N}x{K}")


def test_gemm_residual_ring_with_l2_warm_residual(T):
    """Regression: the three-slot residual ring of the cta_group::2 kernel hands each slot to the two epilogue warpgroups in turn;
    with a parity wait alone a warpgroup could take the OTHER warpgroup's not-yet-landed chunk for its own completed phase (seen
    as flaky launch failures at batch 5, where the residual written by the previous layer is partly L2-resident and TMA latencies
    vary between chunks).  Residual rewritten right before every launch, many launches, every output bit-identical."""
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ops

    torch = T
    M, N, K = 21000, 1024, 256
    a, w = _rand(torch, M, K, seed=51), _rand(torch, N, K, seed=52, scale=K ** -0.5)
    bias, src = torch.randn(N, device="cuda"), _rand(torch, M, N, seed=53)
    res = torch.empty_like(src)
    try:
        _lib.check(_lib.lib().opd_set_option(b"gemm_res_wide", 0), "opd_set_option")
        ref = ops.gemm(a, w, bias, epilogue=2, residual=src)
    finally:
        _lib.lib().opd_set_option(b"gemm_res_wide", 1)
    for it in range(40):
        res.copy_(src)
        got = ops.gemm(a, w, bias, epilogue=2, residual=res)
        assert torch.equal(got.view(torch.int16), ref.view(torch.int16)), f"launch {it}"
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K,epi", [(128 * 7 + 9, 256, 256, 0), (128 * 40, 512, 1024, 1), (128 * 301 + 77, 2048, 256, 1),
                                       (128 * 33, 256, 2048, 2), (128 * 150 + 5, 256, 256, 3), (300, 256, 128, 0)])
@pytest.mark.parametrize("option", [b"gemm_cluster", b"gemm_mpairs", b"gemm_pair"])
def test_gemm_cluster_variant_is_bit_identical(T, M, N, K, epi, option):
    """(gemm_mpairs: the CTA works on pairs of m-blocks that share every weight tile.)  The 2-CTA cluster variant (each CTA loads half of every weight tile and multicasts it to both) against the
    single-CTA kernel: same accumulation order, so the outputs must be bit-identical - even and odd numbers of m-blocks,
    fewer tile pairs than clusters, plain / ReLU / residual / LayerNorm (+pos) epilogues."""
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ops

    torch = T
    a, w = _rand(torch, M, K, seed=21), _rand(torch, N, K, seed=22, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    res = _rand(torch, M, N, seed=23) if epi >= 2 else None
    kw = dict(bias=bias, epilogue=epi, residual=res)
    if epi == 3:
        kw.update(gamma=torch.randn(N, device="cuda"), beta=torch.randn(N, device="cuda"), pos=torch.randn(50, N, device="cuda"))
    try:
        for o in (b"gemm_cluster", b"gemm_mpairs", b"gemm_pair"):
            _lib.check(_lib.lib().opd_set_option(o, 0), "opd_set_option")
        ref = ops.gemm(a, w, **kw)
        _lib.check(_lib.lib().opd_set_option(option, 3 if option == b"gemm_pair" else 2), "opd_set_option")
        got = ops.gemm(a, w, **kw)
        torch.cuda.synchronize()
    finally:
        _lib.lib().opd_set_option(b"gemm_cluster", 0)
        _lib.lib().opd_set_option(b"gemm_mpairs", 0)
        _lib.lib().opd_set_option(b"gemm_pair", PAIR_DEFAULT)
    ref, got = (ref, got) if isinstance(ref, tuple) else ((ref,), (got,))
    for r, g in zip(ref, got):
        assert torch.equal(g.view(torch.int16), r.view(torch.int16))


@pytest.mark.parametrize("B,H,W,C,N,k,stride", [(3, 40, 56, 256, 256, 3, 1), (2, 50, 84, 512, 256, 1, 2)])
@pytest.mark.parametrize("option", [b"gemm_cluster", b"gemm_mpairs", b"gemm_pair"])
def test_conv_cluster_variant_is_bit_identical(T, B, H, W, C, N, k, stride, option):
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ops

    torch = T
    x, w = _rand(torch, B, H, W, C, seed=31), _rand(torch, N, k, k, C, seed=32, scale=(k * k * C) ** -0.5)
    bias = torch.randn(N, device="cuda")
    try:
        for o in (b"gemm_cluster", b"gemm_mpairs", b"gemm_pair"):
            _lib.check(_lib.lib().opd_set_option(o, 0), "opd_set_option")
        ref = ops.conv2d_nhwc(x, w, bias, stride=stride, pad=k // 2, epilogue=1)
        _lib.check(_lib.lib().opd_set_option(option, 3 if option == b"gemm_pair" else 2), "opd_set_option")
        got = ops.conv2d_nhwc(x, w, bias, stride=stride, pad=k // 2, epilogue=1)
        torch.cuda.synchronize()
    finally:
        _lib.lib().opd_set_option(b"gemm_cluster", 0)
        _lib.lib().opd_set_option(b"gemm_mpairs", 0)
        _lib.lib().opd_set_option(b"gemm_pair", PAIR_DEFAULT)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))


@pytest.mark.parametrize("M,K", [(100, 256), (1050 * 3, 256), (777, 2048)])
def test_gemm_layernorm_and_pos(T, M, K):
    from office_person_detection_vit_b200.detection import ops

    torch = T
    N = 256
    a, w = _rand(torch, M, K, seed=4), _rand(torch, N, K, seed=5, scale=K ** -0.5)
    bias, gamma, beta = torch.randn(N, device="cuda"), torch.rand(N, device="cuda") + 0.5, torch.randn(N, device="cuda")
    res = _rand(torch, M, N, seed=6)
    pos_rows = 50 if M % 50 == 0 else M
    pos = torch.randn(pos_rows, N, device="cuda")
    d, d2 = ops.gemm(a, w, bias, epilogue=ops.EPI_BIAS_RES_LN, residual=res, gamma=gamma, beta=beta, pos=pos)
    ref = torch.nn.functional.layer_norm(a.float() @ w.float().T + bias + res.float(), (N,), gamma, beta, 1e-5)
    _close(torch, d, ref, "gemm+LN")
    ref2 = d.float() + pos.repeat(M // pos_rows, 1)
    _close(torch, d2, ref2, "gemm+LN+pos")


@pytest.mark.parametrize("B,H,W,C,N,k,stride,epi", [
    (2, 20, 31, 64, 64, 3, 1, 1),
    (1, 50, 84, 128, 128, 3, 2, 1),
    (3, 25, 42, 256, 256, 3, 2, 1),
    (2, 25, 42, 512, 512, 3, 1, 1),
    (2, 50, 83, 256, 512, 1, 2, 0),
    (2, 17, 23, 64, 256, 1, 1, 2),
    (1, 200, 334, 64, 64, 3, 1, 1),
])
def test_conv_nhwc(T, B, H, W, C, N, k, stride, epi):
    from office_person_detection_vit_b200.detection import ops

    torch = T
    x = _rand(torch, B, H, W, C, seed=7)
    w = _rand(torch, N, k, k, C, seed=8, scale=(C * k * k) ** -0.5)
    bias = torch.randn(N, device="cuda")
    pad = k // 2
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride,
                                     padding=pad).permute(0, 2, 3, 1)
    res = _rand(torch, *ref.shape, seed=9) if epi == 2 else None
    y = ops.conv2d_nhwc(x, w, bias, stride=stride, pad=pad, epilogue=epi, residual=res)
    if epi == 2:
        ref = ref + res.float()
    if epi >= 1:
        ref = ref.relu()
    assert y.shape == ref.shape
    _close(torch, y, ref, f"conv {B}x{H}x{W}x{C}->{N} k{k} s{stride}")


@pytest.mark.parametrize("pair", [1, 0])
@pytest.mark.parametrize("M,with_pos", [(100, True), (128 * 5, False), (1050 * 3 + 17, True), (128 * 149 + 1, True), (128 * 300, True)])
def test_fused_mlp_is_bit_identical_to_two_gemms(T, M, with_pos, pair):
    """tc_mlp.cu (fc1 + ReLU + fc2 + residual + LayerNorm (+ pos) in one kernel, the hidden activations stay on chip) against the
    two GEMM launches it replaces: same accumulation order, same epilogue operations -> the same bits; one tile, ragged last tile,
    more tiles than SMs (CTAs with two tiles); as cta_group::2 pairs (odd tile counts: the last pair repeats a tile) and as single CTAs."""
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ops

    torch = T
    x = _rand(torch, M, 256, seed=61)
    w1, w2 = _rand(torch, 2048, 256, seed=62, scale=256 ** -0.5), _rand(torch, 256, 2048, seed=63, scale=2048 ** -0.5)
    b1, b2 = torch.randn(2048, device="cuda") * 0.1, torch.randn(256, device="cuda") * 0.1
    gamma, beta = torch.rand(256, device="cuda") + 0.5, torch.randn(256, device="cuda") * 0.1
    pos = torch.randn(50, 256, device="cuda") if with_pos else None
    h = ops.gemm(x, w1, b1, epilogue=1)
    ref = ops.gemm(h, w2, b2, epilogue=3, residual=x, gamma=gamma, beta=beta, pos=pos)
    try:
        _lib.check(_lib.lib().opd_set_option(b"mlp_pair", pair), "opd_set_option")
        got = ops.mlp_ln(x, w1, b1, w2, b2, gamma, beta, pos=pos)
        torch.cuda.synchronize()
    finally:
        _lib.lib().opd_set_option(b"mlp_pair", 1)
    ref, got = (ref, got) if with_pos else ((ref,), (got,))
    for r, g in zip(ref, got):
        assert torch.equal(g.view(torch.int16), r.view(torch.int16)), float((g.float() - r.float()).abs().max())
    # and against plain float32 arithmetic
    hf = (x.float() @ w1.float().T + b1).relu().to(torch.bfloat16).float()
    yf = torch.nn.functional.layer_norm(hf @ w2.float().T + b2 + x.float(), (256,), gamma, beta, 1e-5)
    _close(torch, got[0], yf, f"fused mlp M={M}")


@pytest.mark.parametrize("B,Lq,Lk", [(2, 100, 100), (2, 100, 1050), (1, 1050, 1050), (3, 1008, 1008), (1, 7, 65), (2, 128, 256)])
@pytest.mark.parametrize("tc", [4, 1, 2, 3, 0])
def test_attention(T, B, Lq, Lk, tc):
    """tc = 4 / 1 / 2: the tcgen05 / TMEM kernel (attention_tc.cu) with 96 (default) / 64 / 128 keys per tile; tc = 3: 64 keys with
    the S tile held in registers (early S issue); tc = 0: the mma.sync kernel."""
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ops

    torch = T
    D, heads = 256, 8
    q, k, v = _rand(torch, B, Lq, D, seed=10), _rand(torch, B, Lk, D, seed=11), _rand(torch, B, Lk, D, seed=12)
    _lib.check(_lib.lib().opd_set_option(b"attention_tc", int(tc > 0)))
    _lib.check(_lib.lib().opd_set_option(b"attention_kv", {1: 64, 2: 128, 3: 65}.get(tc, 96)))
    try:
        o = ops.attention(q, k, v, heads)
    finally:
        _lib.lib().opd_set_option(b"attention_tc", 1)
        _lib.lib().opd_set_option(b"attention_kv", 96)
    qh, kh, vh = (t.float().view(B, -1, heads, 32).transpose(1, 2) for t in (q, k, v))
    ref = torch.softmax(qh @ kh.transpose(2, 3) * 32 ** -0.5, -1) @ vh
    ref = ref.transpose(1, 2).reshape(B, Lq, D)
    _close(torch, o, ref, f"attention {B}x{Lq}x{Lk}")


@pytest.mark.parametrize("B,H,W,mid,width,stride", [
    (1, 16, 16, 64, 256, 1),
    (2, 50, 83, 64, 256, 1),
    (3, 40, 67, 128, 512, 1),
    (2, 51, 84, 128, 512, 2),
    (1, 200, 334, 64, 256, 1),
    (3, 33, 47, 64, 256, 1),
    (4, 100, 167, 128, 512, 1),
])
@pytest.mark.parametrize("halo", [2, 1, 0])
def test_fused_bottleneck_tail(T, B, H, W, mid, width, stride, halo):
    """conv3x3 + ReLU -> (bf16 in shared memory) -> conv1x1 + bias + residual + ReLU in one kernel; halo = 1 uses the
    halo-patch variant where it applies (64 channels, stride 1), halo = 0 the im2col variant everywhere."""
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ops

    _lib.check(_lib.lib().opd_set_option(b"bneck_halo", halo))

    torch = T
    x = _rand(torch, B, H, W, mid, seed=20)
    w2 = _rand(torch, mid, 3, 3, mid, seed=21, scale=(9 * mid) ** -0.5)
    w3 = _rand(torch, width, mid, seed=22, scale=mid ** -0.5)
    b2, b3 = torch.randn(mid, device="cuda") * 0.3, torch.randn(width, device="cuda") * 0.3
    m = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w2.float().permute(0, 3, 1, 2), b2, stride=stride, padding=1)
    m = m.relu().to(torch.bfloat16).float().permute(0, 2, 3, 1)          # the mid activation is rounded to bf16
    res = _rand(torch, *m.shape[:3], width, seed=23)
    ref = (m @ w3.float().T + b3 + res.float()).relu()
    try:
        y = ops.bottleneck_tail(x, w2, b2, w3, b3, res, stride=stride)
    finally:
        _lib.lib().opd_set_option(b"bneck_halo", 1)
    assert y.shape == ref.shape
    _close(torch, y, ref, f"bottleneck tail {B}x{H}x{W} mid {mid} width {width} s{stride}")


@pytest.mark.parametrize("B,H,W,width,stride", [
    (1, 16, 16, 512, 1),        # one pair of tiles
    (2, 51, 84, 512, 2),        # 18 tiles, the last one partial
    (2, 64, 64, 512, 1),
    (4, 100, 167, 512, 1),      # 522 tiles: several per pair
    (1, 32, 48, 256, 1),
])
def test_fused_bottleneck_tail_pair_variant_is_bit_identical(T, B, H, W, width, stride):
    """cta_group::2 variant of the im2col bottleneck tail (MID = 128, two CTAs = one 256-pixel MMA, half a weight tile per SM):
    same arithmetic in the same order as the one-CTA kernel, so the outputs must be equal bit for bit."""
    from office_person_detection_vit_b200 import _lib
    from office_person_detection_vit_b200.detection import ops

    torch = T
    mid = 128
    x = _rand(torch, B, H, W, mid, seed=30)
    w2 = _rand(torch, mid, 3, 3, mid, seed=31, scale=(9 * mid) ** -0.5)
    w3 = _rand(torch, width, mid, seed=32, scale=mid ** -0.5)
    b2, b3 = torch.randn(mid, device="cuda") * 0.3, torch.randn(width, device="cuda") * 0.3
    P, Q = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    res = _rand(torch, B, P, Q, width, seed=33)
    try:
        _lib.check(_lib.lib().opd_set_option(b"bneck_pair", 0), "opd_set_option")
        a = ops.bottleneck_tail(x, w2, b2, w3, b3, res, stride=stride)
        _lib.check(_lib.lib().opd_set_option(b"bneck_pair", 3), "opd_set_option")
        b = ops.bottleneck_tail(x, w2, b2, w3, b3, res, stride=stride)
        torch.cuda.synchronize()
    finally:
        _lib.lib().opd_set_option(b"bneck_pair", BNECK_PAIR_DEFAULT)
    assert torch.equal(a, b)
