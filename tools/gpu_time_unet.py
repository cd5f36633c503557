"""Times one full-size UNet forward (config-2 shape: B=2, T=16+1, 64x64 latents) through the public class."""
import sys, time
import torch
sys.path.insert(0, ".")
from musev_b200.schema import preset_config
from musev_b200.synth import make_inputs, make_state_dict
from musev_b200.unet import UNet3DConditionModel

dev = "cuda"
preset = sys.argv[1] if len(sys.argv) > 1 else "musev"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 16
hw = int(sys.argv[4]) if len(sys.argv) > 4 else 64
cfg = preset_config(preset)
sd = make_state_dict(cfg, seed=0, dtype=torch.float16)
model = UNet3DConditionModel(cfg, device=dev, dtype=torch.float16)
model.load_state_dict({k: v.to(dev) for k, v in sd.items()})
del sd
inp = make_inputs(cfg, batch=2, frames=frames, h=hw, w=hw, n_vis_cond=1)


def cast(v):
    if torch.is_tensor(v) and v.is_floating_point():
        return v.to(dev, torch.float16)
    if isinstance(v, list):
        return [cast(x) for x in v]
    return v


kw = dict(sample_index=inp["sample_index"], vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"],
          sample_frame_rate=8, down_block_refer_embs=cast(inp.get("down_block_refer_embs")),
          mid_block_refer_emb=cast(inp.get("mid_block_refer_emb")), vision_clip_emb=cast(inp.get("vision_clip_emb")))
x, enc = cast(inp["sample"]), cast(inp["encoder_hidden_states"])
out = model(x, torch.tensor(601), enc, **kw).sample
torch.cuda.synchronize()
print("workspace MB", model._ws.numel() / 2**20, "out std", out.float().std().item(), "nan", torch.isnan(out).any().item(), flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for _ in range(iters):
    out = model(x, torch.tensor(601), enc, **kw).sample
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / max(iters, 1)
print(f"[time] {preset} forward B=2 T={frames}+1 {hw}x{hw}: {ms:.2f} ms/forward (host {1000 * (time.time() - t0) / max(iters, 1):.2f} ms)", flush=True)
