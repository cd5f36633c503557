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
    out_f32: bool = False,
    stride2: bool = False,
) -> torch.Tensor:
    assert a0.dtype == torch.float16 and weight.dtype == torch.float16 and a0.dim() == 4
    assert a0.stride(3) == 1 and weight.is_contiguous()
    NF, H, W, C0 = a0.shape
    N = weight.shape[0]
    nout = N // 2 if geglu else N
    M = NF * H * W if not stride2 else NF * (H // 2) * (W // 2)
    if out is None:
        out = torch.empty((M, nout), dtype=torch.float32 if out_f32 else torch.float16, device=a0.device)
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
    assert stride2 or weight.shape[1] == len(taps) * (C0 + (a1.shape[3] if a1 is not None else 0))
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
    d.out_f32, d.stride2 = int(out_f32), int(stride2)
    _capi.check(_capi.lib().mvb_op_conv_gemm(C.byref(d), _stream()))
    return out


def attention(q, segs, NF, Nq, heads, d, dp, scale, out=None, out_scale=1.0, accumulate=False, v_ones_col=False, variant=0):
    """q: [NF*Nq, >=heads*dp] fp16 (row stride taken from the tensor); segs: list of dicts
    {k, v, nk, fdiv, fmul, fadd} with k/v [rows, >=heads*dp] views sharing a row stride."""
    assert q.dtype == torch.float16 and q.stride(1) == 1
    if out is None:
        out = torch.zeros((NF * Nq, heads * d), dtype=torch.float16, device=q.device)
    a = _capi.AttentionDesc()
    a.q, a.ldq = q.data_ptr(), q.stride(0)
    a.NF, a.Nq, a.heads, a.d, a.dp, a.scale, a.nseg = NF, Nq, heads, d, dp, scale, len(segs)
    for i, s in enumerate(segs):
        k, v = s["k"], s["v"]
        assert k.dtype == torch.float16 and v.dtype == torch.float16 and k.stride(0) == v.stride(0)
        a.k[i], a.v[i], a.ldkv[i], a.kv_rows[i] = k.data_ptr(), v.data_ptr(), k.stride(0), k.shape[0]
        a.nk[i], a.fdiv[i], a.fmul[i], a.fadd[i] = s["nk"], s.get("fdiv", 1), s.get("fmul", s["nk"]), s.get("fadd", 0)
    a.out, a.ldo, a.out_scale, a.accumulate = out.data_ptr(), out.stride(0), out_scale, int(accumulate)
    a.v_ones_col = int(v_ones_col)
    a.variant = int(variant)
    _capi.check(_capi.lib().mvb_op_attention(C.byref(a), _stream()))
    return out


def temporal_attention(qkv, B, T, HW, heads, d, dp, scale):
    assert qkv.dtype == torch.float16 and qkv.is_contiguous()
    out = torch.empty((B * T * HW, heads * d), dtype=torch.float16, device=qkv.device)
    _capi.check(_capi.lib().mvb_op_temporal_attention(qkv.data_ptr(), qkv.shape[-1], B, T, HW, heads, d, dp, scale,
                                                      out.data_ptr(), out.shape[-1], _stream()))
    return out


_gn_barrier = {}     # device -> (zeroed uint32 tensor, ctypes arrival counter) of the one-launch GroupNorm


def groupnorm(x0, gamma, beta, groups=32, frames_per_stat=1, eps=1e-5, silu=False, x1=None, fused=False):
    """x0 [NF, HW, C0] fp16 (+ x1 [NF, HW, C1]); gamma/beta fp32 [C0+C1]. fused=True: the one-launch kernel the engine uses."""
    assert x0.dtype == torch.float16 and x0.is_contiguous() and (x1 is None or (x1.dtype == torch.float16 and x1.is_contiguous()))
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32
    NF, HW, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[2]
    y = torch.empty((NF, HW, C0 + C1), dtype=torch.float16, device=x0.device)
    scratch = torch.empty(NF * 65 * groups * 2, dtype=torch.float32, device=x0.device)
    if fused:
        key = x0.device.index or 0
        if key not in _gn_barrier:
            _gn_barrier[key] = (torch.zeros(1, dtype=torch.int32, device=x0.device), C.c_uint(0))
        word, arrivals = _gn_barrier[key]
        _capi.check(_capi.lib().mvb_op_groupnorm_fused(x0.data_ptr(), C0, _ptr(x1), C1, NF, HW, groups, frames_per_stat, eps,
                                                       gamma.data_ptr(), beta.data_ptr(), int(silu), y.data_ptr(),
                                                       scratch.data_ptr(), word.data_ptr(), C.byref(arrivals), _stream()))
        return y
    _capi.check(_capi.lib().mvb_op_groupnorm(x0.data_ptr(), C0, _ptr(x1), C1, NF, HW, groups, frames_per_stat, eps,
                                             gamma.data_ptr(), beta.data_ptr(), int(silu), y.data_ptr(),
                                             scratch.data_ptr(), _stream()))
    return y


def layernorm(x, gamma, beta, eps):
    assert x.dtype == torch.float16 and x.is_contiguous() and gamma.dtype == torch.float32 and beta.dtype == torch.float32
    M, Cc = x.shape
    y = torch.empty_like(x)
    _capi.check(_capi.lib().mvb_op_layernorm(x.data_ptr(), M, Cc, eps, gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                             _stream()))
    return y


def fuse_cfg_ddim(eps_sum, counter, latents, guidance, alpha_t, alpha_prev, prediction_type=0, clip_range=0.0,
                  out=None, eps_out=None, cfg=True, use_clipped=False, std_dev=0.0, noise=None, x0_out=None):
    """eps_sum fp32 [2B,C,T,H,W] (cfg) or [B,C,T,H,W]; counter fp32 [T] or None; latents fp32/fp16 [B,C,T,H,W]."""
    B, Cc, T = latents.shape[:3]
    HW = latents.shape[3] * latents.shape[4]
    assert eps_sum.dtype == torch.float32 and eps_sum.is_contiguous() and latents.is_contiguous()
    assert latents.dtype in (torch.float16, torch.float32), f"latents must be fp16 or fp32, got {latents.dtype}"
    assert counter is None or (counter.dtype == torch.float32 and counter.is_contiguous() and counter.numel() == T)
    for t_ in (noise, eps_out, x0_out):
        assert t_ is None or (t_.dtype == torch.float32 and t_.is_contiguous())
    if out is None:
        out = torch.empty_like(latents)
    _capi.check(_capi.lib().mvb_fuse_cfg_ddim(
        eps_sum.data_ptr(), _ptr(counter), latents.data_ptr(), out.data_ptr(), int(latents.dtype == torch.float32),
        B, Cc, T, HW, int(cfg), guidance, alpha_t, alpha_prev, prediction_type, clip_range, int(use_clipped),
        std_dev, _ptr(noise), _ptr(eps_out), _ptr(x0_out), _stream()))
    return out


def fuse_cfg_affine(eps_sum, counter, latents, guidance, c_x, c_e, c_n=0.0, noise=None, a_x=0.0, a_e=0.0, aux_out=None,
                    eps_out=None, cfg=True, out=None):
    """x_prev = c_x x + c_e eps + c_n noise (eps = overlap mean + CFG of eps_sum); aux_out = a_x x + a_e eps."""
    B, Cc, T = latents.shape[:3]
    HW = latents.shape[3] * latents.shape[4]
    assert eps_sum.dtype == torch.float32 and eps_sum.is_contiguous() and latents.is_contiguous()
    assert latents.dtype in (torch.float16, torch.float32), f"latents must be fp16 or fp32, got {latents.dtype}"
    assert counter is None or (counter.dtype == torch.float32 and counter.is_contiguous() and counter.numel() == T)
    for t_ in (noise, aux_out, eps_out):
        assert t_ is None or (t_.dtype == torch.float32 and t_.is_contiguous() and t_.numel() == latents.numel())
    if out is None:
        out = torch.empty_like(latents)
    _capi.check(_capi.lib().mvb_fuse_cfg_affine(
        eps_sum.data_ptr(), _ptr(counter), latents.data_ptr(), out.data_ptr(), int(latents.dtype == torch.float32),
        B, Cc, T, HW, int(cfg), guidance, c_x, c_e, c_n, _ptr(noise), a_x, a_e, _ptr(aux_out), _ptr(eps_out), _stream()))
    return out


def accumulate_window(eps_sum, eps_win, src_t0, frames_dev):
    B2, Cc, T = eps_sum.shape[:3]
    HW = eps_sum.shape[3] * eps_sum.shape[4]
    assert eps_sum.dtype == torch.float32 and eps_sum.is_contiguous()
    assert eps_win.dtype in (torch.float16, torch.float32) and eps_win.is_contiguous(), "eps_win: contiguous fp16 / fp32"
    assert frames_dev.dtype == torch.int32 and frames_dev.is_contiguous()
    assert eps_win.shape[0] == B2 and eps_win.shape[1] == Cc and src_t0 + frames_dev.numel() <= eps_win.shape[2]
    _capi.check(_capi.lib().mvb_accumulate_window(eps_sum.data_ptr(), B2, Cc, T, HW, eps_win.data_ptr(),
                                                  int(eps_win.dtype == torch.float32), eps_win.shape[2], src_t0,
                                                  frames_dev.data_ptr(), frames_dev.numel(), _stream()))
