#!/usr/bin/env python
"""Times one UNet forward of a preset at the config-2 shape (CUDA events, N iterations) and prints the per-category split.
Environment knobs (MVB_GN_FUSED, MVB_POLY, MVB_PP_ORDER, ...) are read by the library once per process: A/B = two processes
on the same box."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="musev")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    from musev_b200 import _capi
    from musev_b200.schema import preset_config
    from musev_b200.synth import make_inputs, make_state_dict
    from musev_b200.unet import UNet3DConditionModel
    dev = "cuda"
    cfg = preset_config(a.preset)
    m = UNet3DConditionModel(cfg, device=dev, dtype=torch.float16)
    m.load_state_dict(make_state_dict(cfg, seed=0, dtype=torch.float16))
    inp = make_inputs(cfg, batch=2, frames=16, h=64, w=64, n_vis_cond=1)
    kw = dict(sample_index=inp["sample_index"], vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"], sample_frame_rate=8)
    for k in ("down_block_refer_embs", "mid_block_refer_emb", "vision_clip_emb"):
        if k in inp:
            kw[k] = [x.half().to(dev) for x in inp[k]] if isinstance(inp[k], list) else inp[k].half().to(dev)
    x, enc = inp["sample"].half().to(dev), inp["encoder_hidden_states"].half().to(dev)
    for _ in range(3):
        m(x, 601, enc, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        m(x, 601, enc, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    import time
    host = []
    for _ in range(3):                       # host time to enqueue ONE forward into an empty stream
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m(x, 601, enc, **kw)
        host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    hits, misses = _capi.tensor_map_cache_stats()
    _capi.profile_enable(True)
    m(x, 601, enc, **kw)
    prof = _capi.profile_collect()
    _capi.profile_enable(False)
    print("FWD_TIME " + json.dumps({"tag": a.tag, "preset": a.preset, "ms": ms, "host_enqueue_ms": round(min(host), 2),
                                    "tensor_map_cache": {"hits": hits, "misses": misses},
                                    "split_ms": {k: round(v["ms"], 2) for k, v in prof.items()},
                                    "launches": {k: v["launches"] for k, v in prof.items()}}), flush=True)


if __name__ == "__main__":
    main()
