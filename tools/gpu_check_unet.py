"""GPU bring-up of the whole UNet forward: engine vs the torch-fp32 oracle (run on CUDA) layer by layer.

Usage: python tools/gpu_check_unet.py [preset ...]
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from musev_b200.schema import preset_config
from musev_b200.synth import make_inputs, make_state_dict
from musev_b200.unet import UNet3DConditionModel
from oracle.unet3d_oracle import UNet3DOracle

dev = "cuda"


def to_tokens(x):  # (b t) c h w -> [(b t h w), c]
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])


def run(preset, boc, batch, frames, h, w, dtype, golden=None):
    print(f"=== {preset} boc={boc} B={batch} frames={frames}+1 {h}x{w} io={dtype}", flush=True)
    cfg = preset_config(preset, block_out_channels=boc)
    t0 = time.time()
    sd = make_state_dict(cfg, seed=0)
    print(f"  weights generated in {time.time() - t0:.1f}s", flush=True)
    # the engine stores fp16 weights: give the oracle the same fp16-rounded values
    sd16 = {k: v.half() for k, v in sd.items()}
    inp = make_inputs(cfg, batch=batch, frames=frames, h=h, w=w, n_vis_cond=1)
    t = 601
    model = UNet3DConditionModel(cfg, device=dev, dtype=dtype)
    t0 = time.time()
    model.load_state_dict({k: v.to(dev) for k, v in sd16.items()})
    print(f"  engine loaded in {time.time() - t0:.1f}s", flush=True)
    oracle = UNet3DOracle(cfg, {k: v.float() for k, v in sd16.items()}, device=dev)
    oracle.taps = {}
    kw = dict(sample_index=inp["sample_index"], vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"],
              sample_frame_rate=8, down_block_refer_embs=inp.get("down_block_refer_embs"),
              mid_block_refer_emb=inp.get("mid_block_refer_emb"), vision_clip_emb=inp.get("vision_clip_emb"),
              ip_adapter_scale=0.7)

    def cast(v):
        if torch.is_tensor(v) and v.is_floating_point():
            return v.to(dev, dtype)
        if isinstance(v, list):
            return [cast(x) for x in v]
        return v

    ref = oracle(inp["sample"].to(dtype).float(), t, inp["encoder_hidden_states"].to(dtype).float(),
                 **{k: (cast(v).float() if torch.is_tensor(v) and v.is_floating_point() else
                        ([x.float() for x in cast(v)] if isinstance(v, list) else v)) for k, v in kw.items()})
    out = model(cast(inp["sample"]), torch.tensor(t), cast(inp["encoder_hidden_states"]), do_classifier_free_guidance=True,
                **{k: cast(v) for k, v in kw.items()}).sample
    torch.cuda.synchronize()
    taps = model.debug_taps()
    worst = None
    for name, got in taps.items():
        if name not in oracle.taps:
            continue
        r = to_tokens(oracle.taps[name])
        err = (got - r).abs().max().item()
        scale = r.abs().max().item()
        flag = "" if err <= 0.03 * max(scale, 1.0) else "   <-- DIVERGES"
        print(f"  tap {name:40s} max_abs_err={err:.4e} ref_absmax={scale:.3f}{flag}", flush=True)
        if flag and worst is None:
            worst = name
    err = (out.float() - ref).abs().max().item()
    print(f"  OUTPUT max_abs_err vs oracle(fp32, same fp16 weights) = {err:.4e}  ref std {ref.std().item():.4f} "
          f"nan={torch.isnan(out).any().item()}", flush=True)
    if golden is not None:
        g = torch.load(golden)
        gerr = (out.float().cpu() - g["out"]).abs().max().item()
        oerr = (ref.cpu() - g["out"]).abs().max().item()
        print(f"  vs reference golden ({golden}): engine {gerr:.4e}, oracle {oerr:.4e}", flush=True)
    del model, oracle
    torch.cuda.empty_cache()
    return err


if __name__ == "__main__":
    presets = sys.argv[1:] or ["musev", "musev_referencenet"]
    print(torch.cuda.get_device_name(0), flush=True)
    for p in presets:
        try:
            run(p, (64, 128, 128, 128), 2, 4, 16, 16, torch.float32, golden=f"tests/golden/unet_{p}_narrow.pt")
        except Exception as e:
            import traceback
            traceback.print_exc()
            print("FAILED narrow", p, e, flush=True)
    for p in presets:
        try:
            run(p, (320, 640, 1280, 1280), 2, 2, 8, 8, torch.float32, golden=f"tests/golden/unet_{p}_full.pt")
        except Exception as e:
            import traceback
            traceback.print_exc()
            print("FAILED full", p, e, flush=True)
