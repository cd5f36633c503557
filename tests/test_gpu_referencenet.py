"""ReferenceNet one-shot + IP-Adapter image projection on the engine against the oracle and the
reference golden samples (tests/golden/referencenet_*.pt, produced by the unmodified musev ReferenceNet2D)."""
import os

import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
dev = "cuda"


@pytest.mark.parametrize("tag", ["narrow", "narrow_t2", "full"])
def test_referencenet_vs_oracle_and_reference_golden(built_lib, tag):
    from musev_b200.referencenet import ReferenceNet2D
    from musev_b200.schema import ReferenceNetConfig
    from musev_b200.synth import make_referencenet_inputs, make_state_dict
    from oracle.referencenet_oracle import ReferenceNetOracle
    g = torch.load(os.path.join(GOLDEN, f"referencenet_{tag}.pt"))
    m = g["meta"]
    cfg = ReferenceNetConfig(block_out_channels=tuple(m["block_out_channels"]))
    sd16 = {k: v.half() for k, v in make_state_dict(cfg, seed=m["weight_seed"]).items()}
    model = ReferenceNet2D(cfg, device=dev, dtype=torch.float32)
    model.load_state_dict(sd16)
    oracle = ReferenceNetOracle(cfg, {k: v.float() for k, v in sd16.items()}, device=dev)
    inp = make_referencenet_inputs(cfg, batch=m["batch"], n_ref=m["n_ref"], h=m["h"], w=m["w"], seed=m["input_seed"])
    down, mid, sa = model(inp["sample"].to(dev), torch.zeros((), dtype=torch.long), inp["encoder_hidden_states"].to(dev),
                          num_frames=m["n_ref"], return_ndim=5)
    assert sa is None and len(down) == 12
    rdown, rmid = oracle(inp["sample"], 0, inp["encoder_hidden_states"], num_frames=m["n_ref"], return_ndim=5)
    maps = list(down) + [mid]
    assert [list(x.shape) for x in maps] == m["shapes"]
    for k, (a, b) in enumerate(zip(maps, list(rdown) + [rmid])):
        err = (a.float() - b.float()).abs().max().item()
        assert err < 1e-2 * max(1.0, b.abs().max().item()), f"map {k}: {err}"
    for k, mp in enumerate(maps):                                # the unmodified reference (fp32 weights)
        flat = mp.float().reshape(-1).cpu()
        idx = torch.randint(0, flat.numel(), (m["n_samples"],), generator=torch.Generator().manual_seed(m["sample_seed_base"] + k))
        assert (flat[idx] - g["samples"][k]).abs().max().item() < 1.5e-2 * max(1.0, g["samples"][k].abs().max().item()), k
    # return_ndim=4 keeps (b t) c h w
    d4, m4, _ = model(inp["sample"].to(dev), 0, inp["encoder_hidden_states"].to(dev), num_frames=m["n_ref"], return_ndim=4)
    assert m4.dim() == 4 and torch.equal(m4.view(m["batch"], m["n_ref"], *m4.shape[1:]).permute(0, 2, 1, 3, 4), mid)


def test_referencenet_feeds_the_unet(built_lib):
    """The maps produced on the engine are accepted by the engine's UNet as `down_block_refer_embs` / `mid_block_refer_emb`
    and give the same eps as feeding the oracle's maps (narrow width)."""
    from musev_b200.referencenet import ReferenceNet2D
    from musev_b200.schema import ReferenceNetConfig, preset_config
    from musev_b200.synth import make_inputs, make_referencenet_inputs, make_state_dict
    from musev_b200.unet import UNet3DConditionModel
    from oracle.referencenet_oracle import ReferenceNetOracle
    boc = (64, 128, 128, 128)
    rcfg = ReferenceNetConfig(block_out_channels=boc)
    rsd = {k: v.half() for k, v in make_state_dict(rcfg, seed=5).items()}
    rnet = ReferenceNet2D(rcfg, device=dev, dtype=torch.float32)
    rnet.load_state_dict(rsd)
    rin = make_referencenet_inputs(rcfg, batch=2, n_ref=1, h=16, w=16)
    down, mid, _ = rnet(rin["sample"].to(dev), 0, rin["encoder_hidden_states"].to(dev), num_frames=1)
    rdown, rmid = ReferenceNetOracle(rcfg, {k: v.float() for k, v in rsd.items()}, device=dev)(
        rin["sample"], 0, rin["encoder_hidden_states"], num_frames=1)
    cfg = preset_config("musev_referencenet", block_out_channels=boc)
    unet = UNet3DConditionModel(cfg, device=dev, dtype=torch.float32)
    unet.load_state_dict({k: v.half() for k, v in make_state_dict(cfg, seed=0).items()})
    inp = make_inputs(cfg, batch=2, frames=3, h=16, w=16, n_vis_cond=1, seed=11)
    kw = dict(sample_index=inp["sample_index"], vision_conditon_frames_sample_index=inp["vision_conditon_frames_sample_index"],
              sample_frame_rate=8, vision_clip_emb=inp["vision_clip_emb"].to(dev), ip_adapter_scale=1.0)
    a = unet(inp["sample"].to(dev), 301, inp["encoder_hidden_states"].to(dev), down_block_refer_embs=down,
             mid_block_refer_emb=mid, **kw).sample
    b = unet(inp["sample"].to(dev), 301, inp["encoder_hidden_states"].to(dev), down_block_refer_embs=list(rdown),
             mid_block_refer_emb=rmid, **kw).sample
    assert torch.isfinite(a).all() and (a - b).abs().max().item() < 1e-2


def test_image_proj_and_cfg_grouping(built_lib):
    from musev_b200.referencenet import ImageProjModel, ip_adapter_image_emb
    from musev_b200.schema import ImageProjConfig
    from musev_b200.synth import make_state_dict
    from oracle.referencenet_oracle import image_proj_oracle
    cfg = ImageProjConfig()
    sd = make_state_dict(cfg, seed=9)
    sd16 = {k: (v.half().float() if k == "proj.weight" else v) for k, v in sd.items()}
    proj = ImageProjModel(cfg, device=dev, dtype=torch.float32)
    proj.load_state_dict(sd)
    x = torch.randn(3, 1, 1024, generator=torch.Generator().manual_seed(1))
    got = proj(x.to(dev))
    ref = image_proj_oracle(sd16, x.half().float())
    assert got.shape == (3, 4, 768) and (got.cpu() - ref).abs().max().item() < 5e-3
    emb = ip_adapter_image_emb(proj, x[:2].to(dev), n_images=2, batch_size=1)        # 2 reference images of one video
    assert emb.shape == (2, 8, 768)
    zero = image_proj_oracle(sd16, torch.zeros(2, 1, 1024)).view(1, 8, 768)
    assert (emb[0].cpu() - zero[0]).abs().max().item() < 5e-3                       # uncond = proj(zeros) (:745)
