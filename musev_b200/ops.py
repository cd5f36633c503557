"""Torch-tensor front ends of the op-level C ABI (device pointers + strides are taken from the tensors).

These are thin argument marshalling helpers used by the per-op parity tests; the whole UNet forward runs
inside the library (musev_b200.engine) and does not go through Python per layer.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _capi


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


TAPS_1 = ((0, 0),)
TAPS_3X3 = tuple((dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1))
TAPS_T3 = ((-1, 0), (0, 0), (1, 0))


def conv_gemm(
    a0: torch.Tensor,                      # [NF, H, W, C0] fp16 (any strides with unit channel stride)
    weight: torch.Tensor,                  # [N, ntaps*(C0+C1)] fp16 contiguous
    taps: Sequence[Tuple[int, int]] = TAPS_1,
    a1: Optional[torch.Tensor] = None,
    bias: Optional[torch.Tensor] = None,   # fp32 [N]
    rowadd: Optional[torch.Tensor] = None, # fp32 [groups, N]
    rows_per_group: int = 1,
    residual: Optional[torch.Tensor] = None,  # fp16 [M, Nout]
    alpha: float = 1.0,
    beta: float = 1.0,
    geglu: bool = False,
    act: int = 0,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    assert a0.dtype == torch.float16 and weight.dtype == torch.float16 and a0.dim() == 4
    assert a0.stride(3) == 1 and weight.is_contiguous()
    NF, H, W, C0 = a0.shape
    N = weight.shape[0]
    nout = N // 2 if geglu else N
    M = NF * H * W
    if out is None:
        out = torch.empty((M, nout), dtype=torch.float16, device=a0.device)
    d = _capi.ConvGemmDesc()
    d.a0, d.c0 = a0.data_ptr(), C0
    d.a0_stride_w, d.a0_stride_h, d.a0_stride_n = a0.stride(2), a0.stride(1), a0.stride(0)
    if a1 is not None:
        assert a1.shape[:3] == a0.shape[:3] and a1.stride(3) == 1 and a1.dtype == torch.float16
        d.a1, d.c1 = a1.data_ptr(), a1.shape[3]
        d.a1_stride_w, d.a1_stride_h, d.a1_stride_n = a1.stride(2), a1.stride(1), a1.stride(0)
    d.W, d.H, d.NF = W, H, NF
    d.ntaps = len(taps)
    for i, (dy, dx) in enumerate(taps):
        d.dy[i], d.dx[i] = dy, dx
    assert weight.shape[1] == len(taps) * (C0 + (a1.shape[3] if a1 is not None else 0))
    d.weight, d.N = weight.data_ptr(), N
    d.out, d.ldc = out.data_ptr(), out.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        d.bias = bias.data_ptr()
    if rowadd is not None:
        assert rowadd.dtype == torch.float32 and rowadd.shape[1] == N
        d.rowadd, d.rows_per_group, d.ld_rowadd = rowadd.data_ptr(), rows_per_group, rowadd.stride(0)
    else:
        d.rows_per_group = 1
    if residual is not None:
        assert residual.dtype == torch.float16
        d.residual, d.ld_res = residual.data_ptr(), residual.stride(0)
    d.alpha, d.beta, d.geglu, d.act = alpha, beta, int(geglu), act
    _capi.check(_capi.lib().mvb_op_conv_gemm(C.byref(d), _stream()))
    return out
