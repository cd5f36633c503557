"""CPU tests of the host side: schema, window assignment, scheduler bookkeeping and the multi-rank denoise loop
(world_size 2, gloo) with the device ops replaced by an oracle-backed double."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from conftest import GOLDEN
from musev_b200.context import assign_windows, prepare_global_context
from musev_b200.schema import preset_config, refer_emb_shapes, unet_param_shapes
from musev_b200.synth import make_inputs, make_state_dict


def test_schema_counts_match_reference():
    # 1557 / 1625 tensors, 1419.97 M / 1482.36 M parameters (measured on the reference)
    for preset, n, params in (("musev", 1557, 1419.97e6), ("musev_referencenet", 1625, 1482.36e6)):
        shapes = unet_param_shapes(preset_config(preset))
        assert len(shapes) == n
        total = sum(int(torch.Size(s).numel()) for s in shapes.values())
        assert abs(total - params) / params < 1e-4, total
    shapes, mid = refer_emb_shapes(preset_config("musev_referencenet"), 64, 64)
    assert len(shapes) == 12 and shapes[0] == (320, 64, 64) and shapes[-1] == (1280, 8, 8) and mid == (1280, 8, 8)


def test_preset_errors_like_reference():
    with pytest.raises(ValueError, match="unsupport model_name"):
        preset_config("nope")


@pytest.mark.parametrize("T,win,ov,world", [(48, 16, 4, 4), (128, 16, 4, 8), (512, 16, 8, 8), (16, 16, 4, 8), (48, 16, 4, 3)])
def test_assign_windows_partitions_contiguously(T, win, ov, world):
    ctx = [c[0] for c in prepare_global_context("uniform_v2", 20, T, win, 1, ov, 1)]
    per_rank = assign_windows([len(c) for c in ctx], world)
    flat = [i for r in per_rank for i in r]
    assert flat == list(range(len(ctx)))                      # every window exactly once, contiguous ranges in order
    loads = [sum(len(ctx[i]) + 1 for i in r) for r in per_rank]
    ideal = sum(loads) / world
    assert max(loads) <= ideal + max(len(c) + 1 for c in ctx)  # never worse than one window over the ideal share


def test_scheduler_bookkeeping():
    from musev_b200.scheduler import SD15_DDIM_CONFIG, DDIMScheduler
    from oracle.pipeline_oracle import SD15_DDIM, DDIMOracle
    s = DDIMScheduler(**SD15_DDIM_CONFIG)
    o = DDIMOracle(**SD15_DDIM)
    s.set_timesteps(20)
    o.set_timesteps(20)
    assert s.timesteps.tolist() == o.timesteps.tolist() == [951 - 50 * i for i in range(20)]
    assert s.init_noise_sigma == 1.0 and s.order == 1
    for t in (951, 501, 1):
        a_t, a_p, std = s.step_scalars(t, 0.0)
        assert abs(a_t - float(o.alphas_cumprod[t])) < 1e-7 and std == 0.0
        prev_t = t - 50
        exp = float(o.alphas_cumprod[prev_t]) if prev_t >= 0 else float(o.final_alpha_cumprod)
        assert abs(a_p - exp) < 1e-7
    x = torch.randn(2, 4, 3, 4, 4)
    assert s.scale_model_input(x, 10) is x
    with pytest.raises(ValueError):
        DDIMScheduler().step_scalars(10)                      # set_timesteps not called
    with pytest.raises(ValueError):
        s.set_timesteps(5000)
    import inspect
    params = inspect.signature(s.step).parameters            # the pipeline probes these (pipeline_controlnet.py:1690-1696)
    assert {"eta", "generator", "noise_type", "w_ind_noise"} <= set(params)


# ------------------------------------------------------------------ multi-rank loop (gloo, world_size 2)
class OracleOpsDouble:
    """CPU stand-in for the two device kernels of the loop (tests only; the product uses musev_b200.ops)."""

    @staticmethod
    def accumulate_window(eps_sum, eps, src_t0, frames_dev):
        idx = frames_dev.long()
        eps_sum[:, :, idx] += eps[:, :, src_t0:src_t0 + idx.numel()].float()

    @staticmethod
    def fuse_cfg_ddim(eps_sum, counter, latents, g, a_t, a_p, pred, clip):
        e = eps_sum / counter.view(1, 1, -1, 1, 1)
        u, tx = e.chunk(2)
        e = u + g * (tx - u)
        x0 = (latents - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
        return (a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * e).to(latents.dtype)


class _FakeControlNet:
    """Deterministic stand-in with the diffusers ControlNetModel call signature: residual maps that depend on the sample,
    the timestep, the prompt rows and the condition latents (so that a wrong row / frame slice changes the result)."""

    def __init__(self, shapes):
        self.shapes = shapes      # [(C, downscale)] of the 12 + 1 maps

    def __call__(self, sample, t, enc, controlnet_cond_latents=None, conditioning_scale=1.0, guess_mode=False, return_dict=False):
        nf, _, h, w = sample.shape
        base = (sample.mean(1, keepdim=True) + 0.1 * controlnet_cond_latents.mean(1, keepdim=True)
                + 0.01 * enc.mean((1, 2)).view(nf, 1, 1, 1)) * conditioning_scale * (1.0 + 1e-3 * float(t))
        maps = [torch.nn.functional.avg_pool2d(base, ds).expand(nf, c, h // ds, w // ds).contiguous() * 0.05 if ds > 1
                else base.expand(nf, c, h, w).contiguous() * 0.05 for c, ds in self.shapes]
        return maps[:-1], maps[-1]


def _loop_worker(rank, world, port, preset, out_path, cfg_split=False, with_controlnet=False):
    import torch.distributed as dist
    from musev_b200.pipeline import ParallelDenoiser
    from musev_b200.scheduler import SD15_DDIM_CONFIG, DDIMScheduler
    from oracle.unet3d_oracle import UNet3DOracle
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    g = torch.load(os.path.join(GOLDEN, f"loop_{preset}_narrow.pt"))
    m = g["meta"]
    cfg = preset_config(preset, block_out_channels=tuple(m["block_out_channels"]))
    o = UNet3DOracle(cfg, make_state_dict(cfg, seed=m["weight_seed"]))
    gen = torch.Generator().manual_seed(m["input_seed"])
    latents = torch.randn(1, 4, m["T"], m["h"], m["w"], generator=gen)
    cond = torch.randn(1, 4, 1, m["h"], m["w"], generator=gen) * 0.5
    prompt = torch.randn(2, 77, cfg.cross_attention_dim, generator=gen)
    extra = make_inputs(cfg, batch=2, frames=1, h=m["h"], w=m["w"], seed=m["input_seed"])
    kw = {k: extra[k] for k in ("down_block_refer_embs", "mid_block_refer_emb", "vision_clip_emb") if k in extra}
    kw["ip_adapter_scale"] = 1.0

    def unet(sample, t, enc, return_dict=False, do_classifier_free_guidance=True, **k):
        return (o(sample, t, enc, **k),)

    den = ParallelDenoiser(unet, DDIMScheduler(**SD15_DDIM_CONFIG), device_ops=OracleOpsDouble)
    cn_fn = None
    if with_controlnet:
        from musev_b200.pipeline import make_controlnet_fn
        boc = m["block_out_channels"]
        shapes, ds = [(boc[0], 1)], 1
        for bi, ch in enumerate(boc):
            shapes += [(ch, ds)] * 2
            if bi != len(boc) - 1:
                ds *= 2
                shapes.append((ch, ds))
        shapes.append((boc[-1], ds))
        cn_lat = torch.randn(2, boc[0], 1 + m["T"], m["h"], m["w"], generator=gen)
        cn_fn = make_controlnet_fn(_FakeControlNet(shapes), cn_lat, prompt, 1, controlnet_conditioning_scale=0.7)
    res = den(latents, cond, prompt, num_inference_steps=m["steps"], guidance_scale=m["guidance_scale"],
              context_frames=m["context_frames"], context_overlap=m["context_overlap"], motion_speed=8, unet_kwargs=kw,
              cfg_split=cfg_split, controlnet_fn=cn_fn)
    if rank == 0:
        torch.save({"latents": res.latents, "per_rank": res.windows_per_rank, "windows": res.windows}, out_path)
    if world > 1:
        # replicated state: every rank must hold identical latents
        other = [torch.empty_like(res.latents) for _ in range(world)]
        dist.all_gather(other, res.latents)
        assert all(torch.equal(other[0], x) for x in other)
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("preset", ["musev"])
def test_parallel_denoise_two_ranks_gloo(tmp_path, preset):
    g = torch.load(os.path.join(GOLDEN, f"loop_{preset}_narrow.pt"))
    p1 = str(tmp_path / "w1.pt")
    _loop_worker(0, 1, 0, preset, p1)
    single = torch.load(p1)
    assert single["windows"] == g["meta"]["contexts"]
    assert (single["latents"] - g["latents"]).abs().max().item() < 2e-4      # host loop == reference loop
    p2 = str(tmp_path / "w2.pt")
    mp.spawn(_loop_worker, args=(2, _free_port(), preset, p2), nprocs=2, join=True)
    double = torch.load(p2)
    assert sorted(i for r in double["per_rank"] for i in r) == list(range(len(double["windows"])))
    assert all(len(r) >= 1 for r in double["per_rank"])
    # sharding changes only the summation order of the overlap accumulation
    assert (double["latents"] - single["latents"]).abs().max().item() < 1e-5
    # CFG split on the same two ranks: one pair, rank 0 = unconditional half, rank 1 = text half of every window
    p3 = str(tmp_path / "w3.pt")
    mp.spawn(_loop_worker, args=(2, _free_port(), preset, p3, True), nprocs=2, join=True)
    split = torch.load(p3)
    assert split["per_rank"] == [list(range(len(split["windows"])))]
    # a B-row forward and a 2B-row forward of the CPU oracle differ by fp32 blocking order (measured 2e-5)
    assert (split["latents"] - single["latents"]).abs().max().item() < 1e-4
    # with a ControlNet callback: window slicing of the condition latents and, under CFG split, of the prompt / batch rows
    p4, p5 = str(tmp_path / "w4.pt"), str(tmp_path / "w5.pt")
    _loop_worker(0, 1, 0, preset, p4, False, True)
    mp.spawn(_loop_worker, args=(2, _free_port(), preset, p5, True, True), nprocs=2, join=True)
    cn1, cn2 = torch.load(p4), torch.load(p5)
    assert (cn1["latents"] - single["latents"]).abs().max().item() > 1e-3      # the residuals matter
    assert (cn2["latents"] - cn1["latents"]).abs().max().item() < 1e-4


def test_closed_loop_window_with_repeated_frames_matches_reference_semantics():
    """`uniform` with context_stride 2 and context_frames < T < 2 context_frames wraps (e % num_frames, context.py:46) and
    names frames twice inside one window. The reference's indexed assignment keeps the LAST occurrence and bumps the
    counter once (pipeline_controlnet.py:2076-2077); the loop must do the same (it used to add both and count twice)."""
    from musev_b200.pipeline import ParallelDenoiser
    from musev_b200.scheduler import SD15_DDIM_CONFIG, DDIMScheduler
    from oracle.pipeline_oracle import SD15_DDIM, DDIMOracle, denoise_loop
    T, h, w = 20, 4, 4
    ctx = [c[0] for c in prepare_global_context("uniform", 2, T, 12, 2, 4, 1)]
    assert any(len(set(c)) < len(c) for c in ctx), "this schedule is expected to repeat frames inside a window"
    g = torch.Generator().manual_seed(3)
    latents = torch.randn(1, 4, T, h, w, generator=g)
    cond = torch.randn(1, 4, 1, h, w, generator=g)
    prompt = torch.randn(2, 77, 8, generator=g)

    def fake_unet(sample, t, enc, **k):
        # frame-position dependent, so that the two occurrences of a repeated frame give different eps
        pos = torch.arange(sample.shape[2], dtype=sample.dtype).view(1, 1, -1, 1, 1)
        return torch.sin(sample * 1.3 + 0.01 * float(t) + 0.37 * pos) + enc.mean() * 0.1

    den = ParallelDenoiser(lambda s, t, e, return_dict=False, do_classifier_free_guidance=True, **k: (fake_unet(s, t, e),),
                           DDIMScheduler(**SD15_DDIM_CONFIG), device_ops=OracleOpsDouble)
    res = den(latents, cond, prompt, num_inference_steps=3, guidance_scale=2.0, context_frames=12, context_overlap=4,
              context_schedule="uniform", context_stride=2)
    ref = denoise_loop(fake_unet, DDIMOracle(**SD15_DDIM), latents, cond, prompt, 3, 2.0, context_frames=12, context_overlap=4,
                       context_schedule="uniform", context_stride=2)
    assert res.windows == ctx
    assert (res.latents - ref).abs().max().item() < 1e-5


def test_sampler_mirrors_bookkeeping_and_affine_scalars():
    """EulerDiscreteScheduler / LCMScheduler host mirrors: timesteps, sigmas, init_noise_sigma equal the oracle's, and the
    affine scalars reproduce the oracle's (reference-form) step on CPU tensors."""
    import inspect
    from musev_b200.samplers import EulerDiscreteScheduler, LCMScheduler
    from oracle.sampler_oracle import EulerOracle, LCMOracle
    kw = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
              timestep_spacing="leading", steps_offset=1)
    for pred in ("epsilon", "v_prediction"):
        s, o = EulerDiscreteScheduler(prediction_type=pred, **kw), EulerOracle(prediction_type=pred, **kw)
        s.set_timesteps(20)
        o.set_timesteps(20)
        assert torch.equal(s.timesteps, o.timesteps) and torch.allclose(s.sigmas, o.sigmas)
        assert abs(float(s.init_noise_sigma) - float(o.init_noise_sigma)) < 1e-6
        g = torch.Generator().manual_seed(0)
        x, e = torch.randn(2, 4, 3, 4, 4, generator=g), torch.randn(2, 4, 3, 4, 4, generator=g)
        for t in s.timesteps[:5]:
            assert abs(s.model_input_scale(t) - float(o.scale_model_input(torch.ones(1), t))) < 1e-6
            a = s.affine_step(t)
            prev, x0 = o.step(e, t, x)
            assert (a.c_x * x + a.c_e * e - prev).abs().max().item() < 2e-5 and a.c_n == 0.0
            assert (a.a_x * x + a.a_e * e - x0).abs().max().item() < 2e-5
    with pytest.raises(ValueError, match="integer indices"):
        s.step(e, 3, x)
    assert {"generator", "noise_type", "w_ind_noise", "s_churn"} <= set(inspect.signature(s.step).parameters)
    l, lo = LCMScheduler(), LCMOracle()
    l.set_timesteps(4)
    lo.set_timesteps(4)
    assert l.timesteps.tolist() == lo.timesteps.tolist() == [999, 759, 499, 259]
    gen = torch.Generator().manual_seed(3)
    for i, t in enumerate(l.timesteps):
        a = l.affine_step(t)
        prev, den, noise = lo.step(e, t, x, gen)
        mine = a.c_x * x + a.c_e * e + (a.c_n * noise if noise is not None else 0)
        assert (mine - prev).abs().max().item() < 2e-5 and (a.a_x * x + a.a_e * e - den).abs().max().item() < 2e-5
        assert (a.c_n == 0.0) == (i == 3)
    with pytest.raises(NotImplementedError):
        LCMScheduler(clip_sample=True)
