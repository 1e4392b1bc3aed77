"""ctypes wrappers of the tensor-core building blocks (include/opd_b200.h: opd_gemm_bf16, opd_conv2d_nhwc_bf16,
opd_attention_bf16).  Torch tensors are used as device buffers only; every call enqueues hand-written sm_100a
kernels on the current stream.  These are the unit-test surface of csrc/tc_gemm.cu and csrc/attention.cu."""

from __future__ import annotations

import ctypes as C

from .. import _lib

_P = C.c_void_p
_lib.register("opd_gemm_bf16", C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        _P, _P, C.c_int64, _P, _P, _P, _P, C.c_int32, _P])
_lib.register("opd_conv2d_nhwc_bf16", C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P])
_lib.register("opd_attention_bf16", C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, _P])

_lib.register("opd_bottleneck_tail_bf16", C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P,
                                                   C.c_int32, _P, _P, _P])

_lib.register("opd_mlp_ln_bf16", C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P])

EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_RES_RELU, EPI_BIAS_RES_LN = 0, 1, 2, 3


def mlp_ln(x, w1, b1, w2, b2, gamma, beta, pos=None):
    """LayerNorm(relu(x w1^T + b1) w2^T + b2 + x) * gamma + beta in one kernel (and D2 = D + pos[row % len(pos)] if pos is given)."""
    torch = _lib.require_cuda()
    M = x.shape[0]
    d = torch.empty_like(x)
    d2 = torch.empty_like(x) if pos is not None else None
    rc = _lib.lib().opd_mlp_ln_bf16(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(gamma), _lib.ptr(beta),
                                    _lib.ptr(d), _lib.ptr(d2), _lib.ptr(pos), pos.shape[0] if pos is not None else 0, M,
                                    _lib.stream_ptr())
    _lib.check(rc, "opd_mlp_ln_bf16")
    return (d, d2) if pos is not None else d


def gemm(a, w, bias=None, epilogue=EPI_BIAS, residual=None, gamma=None, beta=None, pos=None):
    """D = epilogue(a @ w.T + bias); a [M,K] bf16, w [N,K] bf16 -> D [M,N] bf16 (and D2 = D + pos if pos is given)."""
    torch = _lib.require_cuda()
    M, K = a.shape
    N = w.shape[0]
    d = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    d2 = torch.empty_like(d) if pos is not None else None
    rc = _lib.lib().opd_gemm_bf16(_lib.ptr(a), a.stride(0), _lib.ptr(w), _lib.ptr(d), N, M, N, K, epilogue,
                                  _lib.ptr(bias), _lib.ptr(residual), residual.stride(0) if residual is not None else 0,
                                  _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(d2), _lib.ptr(pos),
                                  pos.shape[0] if pos is not None else 0, _lib.stream_ptr())
    _lib.check(rc, "opd_gemm_bf16")
    return (d, d2) if pos is not None else d


def conv2d_nhwc(x, w, bias=None, stride=1, pad=0, epilogue=EPI_BIAS, residual=None):
    """x [B,H,W,C] bf16, w [N,KH,KW,C] bf16 -> y [B,P,Q,N] bf16."""
    torch = _lib.require_cuda()
    B, H, W, Cc = x.shape
    N, KH, KW, _ = w.shape
    P, Q = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    y = torch.empty(B, P, Q, N, dtype=torch.bfloat16, device=x.device)
    rc = _lib.lib().opd_conv2d_nhwc_bf16(_lib.ptr(x), B, H, W, Cc, _lib.ptr(w), N, KH, KW, stride, pad, epilogue,
                                         _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y), _lib.stream_ptr())
    _lib.check(rc, "opd_conv2d_nhwc_bf16")
    return y


def attention(q, k, v, heads=8):
    """q [B,Lq,D], k/v [B,Lk,D] bf16 (last dim contiguous, batch stride = L * row stride) -> o [B,Lq,D] bf16."""
    torch = _lib.require_cuda()
    B, Lq, D = q.shape
    Lk = k.shape[1]
    o = torch.empty(B, Lq, D, dtype=torch.bfloat16, device=q.device)
    rc = _lib.lib().opd_attention_bf16(_lib.ptr(q), q.stride(1), _lib.ptr(k), k.stride(1), _lib.ptr(v), v.stride(1),
                                       _lib.ptr(o), D, B, heads, Lq, Lk, _lib.stream_ptr())
    _lib.check(rc, "opd_attention_bf16")
    return o


def bottleneck_tail(x, w2, bias2, w3, bias3, residual, stride=1):
    """y = relu(conv1x1(relu(conv3x3(x, stride) + bias2), w3) + bias3 + residual), one fused kernel.
    x [B,H,W,mid], w2 [mid,3,3,mid], w3 [width,mid], residual [B,P,Q,width] (all bf16) -> y like residual."""
    torch = _lib.require_cuda()
    B, H, W, mid = x.shape
    width = w3.shape[0]
    y = torch.empty_like(residual)
    rc = _lib.lib().opd_bottleneck_tail_bf16(_lib.ptr(x), B, H, W, mid, _lib.ptr(w2), _lib.ptr(bias2), stride, _lib.ptr(w3),
                                             _lib.ptr(bias3), width, _lib.ptr(residual), _lib.ptr(y), _lib.stream_ptr())
    _lib.check(rc, "opd_bottleneck_tail_bf16")
    return y
