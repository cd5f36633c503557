"""Host mirror of diffusers `ControlNetModel` as MuseV runs it per window-step.

Reference: diffusers/src/diffusers/models/controlnet.py:645-852, called from
musev/pipelines/pipeline_controlnet.py:1238-1262. The SD-1.5 encoder half, the 12 + 1 zero convolutions and the output
scaling run inside libmusevb200.so (`mvb_controlnet_forward`, musev_b200/csrc/engine.cu). The conditioning embedding
(controlnet.py:101-112) is a one-shot conv stack on the 8x larger condition image; the pipeline computes it once per
call and passes `controlnet_cond_latents` on every step (pipeline_controlnet.py:1258), so it stays a handful of torch
convolutions here, outside the per-step path.

"""
from __future__ import annotations

import ctypes as C
from dataclasses import asdict
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.nn.functional as F

from . import _capi
from .schema import ControlNetConfig, controlnet_param_shapes
from .unet import MvbConfig, _is_f32, _lib as _unet_lib, load_weights_batched

MAX_OUT = 13


class MvbControlnetArgs(C.Structure):
    _fields_ = [
        ("sample", C.c_void_p), ("sample_is_f32", C.c_int),
        ("NF", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("timestep", C.c_float),
        ("encoder_hidden_states", C.c_void_p), ("ehs_is_f32", C.c_int), ("n_text", C.c_int),
        ("cond_latents", C.c_void_p), ("cond_is_f32", C.c_int),
        ("n_out", C.c_int),
        ("scales", C.c_float * MAX_OUT),
        ("outs", C.c_void_p * MAX_OUT),
        ("out_is_f32", C.c_int),
        ("out_frames", C.c_int),
    ]


_declared = False


def _lib():
    global _declared
    l = _unet_lib()
    if not _declared:
        l.mvb_create_controlnet.argtypes = [C.POINTER(MvbConfig), C.c_int, C.POINTER(C.c_void_p)]
        l.mvb_create_controlnet.restype = C.c_int
        l.mvb_controlnet_workspace_bytes.argtypes = [C.c_void_p, C.POINTER(MvbControlnetArgs)]
        l.mvb_controlnet_workspace_bytes.restype = C.c_longlong
        l.mvb_controlnet_forward.argtypes = [C.c_void_p, C.POINTER(MvbControlnetArgs), C.c_void_p, C.c_longlong, C.c_void_p]
        l.mvb_controlnet_forward.restype = C.c_int
        _declared = True
    return l


class ControlNetOutput(SimpleNamespace):
    """diffusers models/controlnet.py:46-61."""


class ControlNetModel:
    """B200 engine behind the call surface of diffusers `ControlNetModel` (models/controlnet.py:114)."""

    def __init__(self, config: ControlNetConfig, device: Union[str, torch.device] = "cuda", dtype: torch.dtype = torch.float16):
        if not torch.cuda.is_available():
            raise RuntimeError("musev_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        self.cfg = config
        self.device = torch.device(device if str(device) != "cuda" else f"cuda:{torch.cuda.current_device()}")
        self.dtype = dtype
        self.config = SimpleNamespace(**asdict(config), global_pool_conditions=False)
        self._ws: Optional[torch.Tensor] = None
        self._h = C.c_void_p()
        self._loaded = False
        self._cond_w: Dict[str, torch.Tensor] = {}
        c = MvbConfig()
        c.in_channels, c.out_channels = config.in_channels, config.in_channels
        c.num_blocks = len(config.block_out_channels)
        for i, v in enumerate(config.block_out_channels):
            c.block_out_channels[i] = v
        c.layers_per_block, c.heads = config.layers_per_block, config.attention_head_dim
        c.cross_attention_dim, c.norm_num_groups, c.norm_eps = config.cross_attention_dim, config.norm_num_groups, config.norm_eps
        rc = _lib().mvb_create_controlnet(C.byref(c), self.device.index or 0, C.byref(self._h))
        if rc != 0:
            raise _capi.MvbError(f"mvb_create_controlnet failed ({rc}): unsupported configuration or out of device memory")
        # residual map geometry: (channels, downscale) of the 12 + 1 outputs (controlnet.py:788-823)
        self._maps: List[Tuple[int, int]] = [(config.block_out_channels[0], 1)]
        ds = 1
        nb = len(config.block_out_channels)
        for i, ch in enumerate(config.block_out_channels):
            for _ in range(config.layers_per_block):
                self._maps.append((ch, ds))
            if i != nb - 1:
                ds *= 2
                self._maps.append((ch, ds))
        self._maps.append((config.block_out_channels[-1], ds))

    @classmethod
    def from_state_dict(cls, state_dict: Dict[str, torch.Tensor], device="cuda", dtype=torch.float16,
                        **config_overrides) -> "ControlNetModel":
        m = cls(ControlNetConfig(**config_overrides), device=device, dtype=dtype)
        m.load_state_dict(state_dict)
        return m

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        expected = controlnet_param_shapes(self.cfg)
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        l = _lib()
        todo = []
        for name, shape in expected.items():
            if name not in state_dict:
                continue
            t = state_dict[name]
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError(f"size mismatch for {name}: {tuple(t.shape)} vs {tuple(shape)}")
            if name.startswith("controlnet_cond_embedding."):
                self._cond_w[name] = t.to(self.device, self.dtype).contiguous()
                continue
            todo.append((name, t))
        load_weights_batched(self._h, todo, self.device)
        rc = l.mvb_finalize(self._h)
        if rc != 0:
            raise _capi.MvbError(f"mvb_finalize: {l.mvb_handle_error(self._h).decode()}")
        self._loaded = True
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _lib().mvb_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def eval(self):
        return self

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype):
                if a not in (torch.float16, torch.float32):
                    raise ValueError("musev_b200 computes in fp16 with fp32 accumulation; I/O dtype is fp16 or fp32")
                self.dtype = a
            elif isinstance(a, (str, torch.device)) and torch.device(a).type != "cuda":
                raise RuntimeError("musev_b200 has no CPU path")
        return self

    # ------------------------------------------------------------------ one-shot conditioning embedding
    @torch.no_grad()
    def controlnet_cond_embedding(self, conditioning: torch.Tensor) -> torch.Tensor:
        """ControlNetConditioningEmbedding.forward (models/controlnet.py:101-112); once per pipeline call."""
        p = "controlnet_cond_embedding"
        w = self._cond_w
        x = conditioning.to(self.device, self.dtype)
        e = F.silu(F.conv2d(x, w[p + ".conv_in.weight"], w[p + ".conv_in.bias"], padding=1))
        n_blocks = 2 * (len(self.cfg.conditioning_embedding_out_channels) - 1)
        for i in range(n_blocks):
            e = F.silu(F.conv2d(e, w[f"{p}.blocks.{i}.weight"], w[f"{p}.blocks.{i}.bias"], padding=1, stride=2 if i % 2 else 1))
        return F.conv2d(e, w[p + ".conv_out.weight"], w[p + ".conv_out.bias"], padding=1)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(
        self,
        sample: torch.Tensor,
        timestep: Union[torch.Tensor, float, int],
        encoder_hidden_states: torch.Tensor,
        controlnet_cond: Optional[torch.Tensor] = None,
        conditioning_scale: float = 1.0,
        class_labels: Optional[torch.Tensor] = None,
        timestep_cond: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        added_cond_kwargs: Optional[Dict[str, torch.Tensor]] = None,
        cross_attention_kwargs: Optional[Dict[str, Any]] = None,
        guess_mode: bool = False,
        return_dict: bool = True,
        controlnet_cond_latents: Optional[torch.Tensor] = None,
    ):
        """Reference: ControlNetModel.forward, diffusers models/controlnet.py:645-852."""
        if not self._loaded:
            raise RuntimeError("weights not loaded: call load_state_dict first")
        for name, v in (("class_labels", class_labels), ("timestep_cond", timestep_cond), ("attention_mask", attention_mask),
                        ("added_cond_kwargs", added_cond_kwargs)):
            if v is not None:
                raise NotImplementedError(f"{name} is not used by the SD-1.5 ControlNets MuseV loads")
        if sample.dim() != 4:
            raise ValueError(f"sample must be (b t) c h w, got {tuple(sample.shape)}")
        if controlnet_cond_latents is None:
            if controlnet_cond is None:
                raise ValueError("controlnet_cond or controlnet_cond_latents is required")
            controlnet_cond_latents = self.controlnet_cond_embedding(controlnet_cond)
        NF, _, H, W = sample.shape
        if encoder_hidden_states.dim() != 3 or encoder_hidden_states.shape[0] != NF:
            raise ValueError("encoder_hidden_states must be [(b t), n_text, dim] (one row block per frame)")
        dev = self.device
        sample = sample.to(dev).contiguous()
        ehs = encoder_hidden_states.to(dev).contiguous()
        cond = controlnet_cond_latents.to(dev).contiguous()
        if tuple(cond.shape) != (NF, self.cfg.block_out_channels[0], H, W):
            raise ValueError(f"controlnet_cond_latents has shape {tuple(cond.shape)}")
        t_val = float(timestep.reshape(-1)[0].item()) if torch.is_tensor(timestep) else float(timestep)
        n_out = len(self._maps)
        scale = float(conditioning_scale)
        if guess_mode:   # :826-830
            scales = (torch.logspace(-1, 0, n_out) * scale).tolist()
        else:            # :831-833
            scales = [scale] * n_out
        outs = [torch.empty((NF, c, H // ds, W // ds), device=dev, dtype=self.dtype) for c, ds in self._maps]
        a = MvbControlnetArgs()
        a.sample, a.sample_is_f32 = sample.data_ptr(), _is_f32(sample)
        a.NF, a.H, a.W = NF, H, W
        a.timestep = t_val
        a.encoder_hidden_states, a.ehs_is_f32, a.n_text = ehs.data_ptr(), _is_f32(ehs), ehs.shape[1]
        a.cond_latents, a.cond_is_f32 = cond.data_ptr(), _is_f32(cond)
        a.n_out = n_out
        for k in range(n_out):
            a.scales[k] = scales[k]
            a.outs[k] = outs[k].data_ptr()
        a.out_is_f32 = _is_f32(outs[0])
        l = _lib()
        need = l.mvb_controlnet_workspace_bytes(self._h, C.byref(a))
        if need < 0:
            raise _capi.MvbError(f"mvb_controlnet_workspace_bytes: {l.mvb_handle_error(self._h).decode()}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        rc = l.mvb_controlnet_forward(self._h, C.byref(a), self._ws.data_ptr(), self._ws.numel(),
                                      torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise _capi.MvbError(f"mvb_controlnet_forward: {l.mvb_handle_error(self._h).decode()}")
        down, mid = outs[:-1], outs[-1]
        if not return_dict:
            return (down, mid)
        return ControlNetOutput(down_block_res_samples=down, mid_block_res_sample=mid)

    __call__ = forward
