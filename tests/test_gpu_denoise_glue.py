"""-m gpu: SURVEY.md §8f row N1 — guidance combine + CogVideoXDPMScheduler.step + model-input write as one kernel, and
the device-resident denoising loop, against oracle/dpm_oracle.py (the pipeline's own torch expressions).  Bit-exact:
every product / sum is rounded where torch's promotion rules round it."""
import dataclasses

import pytest
import torch

pytestmark = pytest.mark.gpu


def _same(a, b, tag):
    a, b = a.detach().cpu(), b.detach().cpu()
    if not torch.equal(a, b):
        bad = a != b
        raise AssertionError(f"{tag}: {int(bad.sum())} of {a.numel()} elements differ, max |diff| "
                             f"{float((a.float() - b.float())[bad].abs().max())}, first {a[bad][:3].tolist()} vs "
                             f"{b[bad][:3].tolist()}, nan {int(a.float().isnan().sum())}/{int(b.float().isnan().sum())}")


def _coef_rows(sch, ts, guidance):
    return [sch.step_coefficients(t, ts[i - 1] if i else None, i > 0, guidance) for i, t in enumerate(ts)]


@pytest.mark.parametrize("prediction_type", ["v_prediction", "epsilon", "sample"])
@pytest.mark.parametrize("cfg_batch", [2, 1])
def test_cfg_dpm_step_kernel_bit_exact(built, prediction_type, cfg_batch):
    import bya_b200  # noqa: F401
    from bya_b200 import ops
    from bya_b200.scheduler import _PRED, CogVideoXDPMScheduler
    from oracle.dpm_oracle import DPMSchedulerOracle

    F, C, H, W, Cin = 3, 16, 10, 14, 48
    n = F * C * H * W
    # epsilon prediction divides by sqrt(alpha_t): keep alpha_999 > 0 for it (CogVideoX itself is v_prediction)
    kw = dict(prediction_type=prediction_type, rescale_betas_zero_snr=prediction_type != "epsilon")
    sch, orc = CogVideoXDPMScheduler.cogvideox_5b(**kw), DPMSchedulerOracle(**kw)
    sch.set_timesteps(5)   # 999, 799, ..., 199: the last step lands on alpha = 1 (first order, no noise)
    orc.set_timesteps(5)
    ts = sch.timesteps.tolist()
    g = 6.0 if cfg_batch == 2 else 1.0
    coef = torch.tensor(_coef_rows(sch, ts, g), dtype=torch.float32, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(5)
    noise = torch.randn(len(ts), 2, n, generator=gen, device="cuda", dtype=torch.bfloat16)
    idx = torch.zeros(1, dtype=torch.int32, device="cuda")
    x = torch.randn(1, F, C, H, W, generator=gen, device="cuda", dtype=torch.bfloat16)
    old = torch.zeros(1, F, C, H, W, device="cuda", dtype=torch.float32)
    model_in = torch.full((cfg_batch, F, Cin, H, W), 7.0, device="cuda", dtype=torch.bfloat16)
    x_ref, old_ref = x.clone(), None
    for i, t in enumerate(ts):   # first-order first step, second-order middle steps, first-order last step
        out = 0.5 * torch.randn(cfg_batch, F, C, H, W, generator=gen, device="cuda", dtype=torch.bfloat16)
        idx.fill_(i)
        ops.cfg_dpm_step(out, x, x, old, old, noise, coef, prediction_type=_PRED[prediction_type], step_index=idx,
                         model_input=model_in)
        for device in ("cuda", "cpu"):   # the reference evaluates the expression on the device; the CPU must agree too
            o = out.float().to(device)
            if cfg_batch == 2:
                u, c = o.chunk(2)
                o = u + g * (c - u)

            def run(**k):
                draws = iter([noise[i, 0], noise[i, 1]])
                return orc.step(o, None if old_ref is None else old_ref.to(device), t, ts[i - 1] if i else None,
                                x_ref.to(device), lambda shape, dtype: next(draws).reshape(shape).to(device), **k)

            prev, pred = run()
            prev = prev.to(torch.bfloat16)
            _same(pred, old, f"pred_original_sample {prediction_type} B={cfg_batch} step {i} on {device}")
            _same(prev, x, f"prev_sample {prediction_type} B={cfg_batch} step {i} on {device}")
            if device == "cuda":   # the published expressions as written, evaluated by torch on the device (= the reference)
                prev_n, pred_n = run(device_semantics=False)
                _same(pred_n, pred, f"native pred {prediction_type} step {i}")
                _same(prev_n.to(torch.bfloat16), prev, f"native prev {prediction_type} step {i}")
        x_ref, old_ref = prev.cuda(), pred.cuda()
        assert torch.equal(model_in[:, :, :C], x.expand(cfg_batch, -1, -1, -1, -1))
        assert bool((model_in[:, :, C:] == 7.0).all())
    assert torch.isfinite(x.float()).all()
    assert coef[0, 8] == 0 and coef[-1, 8] == 0 and bool((coef[1:-1, 8] == 1).all())


def test_cfg_dpm_step_rejects_bad_arguments(built):
    import bya_b200  # noqa: F401
    from bya_b200 import ops

    n = 2 * 16 * 4 * 6
    x = torch.zeros(1, 2, 16, 4, 6, device="cuda", dtype=torch.bfloat16)
    p = torch.zeros(1, 2, 16, 4, 6, device="cuda", dtype=torch.float32)
    noise = torch.zeros(1, 2, n, device="cuda", dtype=torch.bfloat16)
    coef = torch.zeros(1, 12, device="cuda", dtype=torch.float32)
    ops.cfg_dpm_step(x, x, x, p, p, noise, coef)
    with pytest.raises(RuntimeError):
        ops.cfg_dpm_step(x, x, x, p, p, noise, torch.zeros(2, 12, device="cuda"))            # rows without step_index
    with pytest.raises(RuntimeError):
        ops.cfg_dpm_step(x, x, x, p.half(), p, noise, coef)                                  # old_pred dtype
    with pytest.raises(RuntimeError):
        ops.cfg_dpm_step(torch.cat([x, x, x]), x, x, p, p, noise, coef)                      # three branches
    with pytest.raises(RuntimeError):
        ops.cfg_dpm_step(x, x, x, p, p, noise, coef, prediction_type=7)
    with pytest.raises(RuntimeError):
        ops.cfg_dpm_step(x, x, x, p, p, noise, coef, model_input=torch.zeros(1, 2, 8, 4, 6, device="cuda", dtype=torch.bfloat16))
    odd = torch.zeros(1, 1, 3, 1, 1, device="cuda", dtype=torch.bfloat16)                    # 3 elements: not 8-divisible
    with pytest.raises(RuntimeError):
        ops.cfg_dpm_step(odd, odd, odd, odd.float(), odd.float(), torch.zeros(1, 2, 3, device="cuda", dtype=torch.bfloat16), coef)


def test_select_step_kernel(built):
    import bya_b200  # noqa: F401
    from bya_b200 import ops

    ts = torch.tensor([999, 500, 19], device="cuda")
    t = torch.zeros(2, dtype=torch.int64, device="cuda")
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    idx = torch.full((1,), -5, dtype=torch.int32, device="cuda")
    for i in range(4):   # the fourth call clamps to the last step instead of reading past the table
        ops.denoise_select_step(ts, t, counter, idx)
        assert t.tolist() == [int(ts[min(i, 2)])] * 2 and int(idx) == min(i, 2)


def test_scheduler_step_drop_in_same_seed_same_latents(built):
    """`scheduler.step(noise_pred, old_pred, t, t_back, latents, generator=...)` the way the pipeline calls it (:934-943)."""
    import bya_b200  # noqa: F401
    from bya_b200.scheduler import CogVideoXDPMScheduler
    from oracle.dpm_oracle import DPMSchedulerOracle

    sch, orc = CogVideoXDPMScheduler.cogvideox_5b(), DPMSchedulerOracle()
    sch.set_timesteps(5, device="cuda")
    orc.set_timesteps(5)
    assert sch.timesteps.is_cuda
    g1, g2 = torch.Generator(device="cuda").manual_seed(11), torch.Generator(device="cuda").manual_seed(11)
    lat = torch.randn(1, 4, 16, 6, 8, device="cuda", dtype=torch.bfloat16)
    lat_ref, old, old_ref = lat.clone(), None, None
    for i, t in enumerate(sch.timesteps):
        npred = torch.randn(1, 4, 16, 6, 8, device="cuda")   # fp32, after .float() and the guidance combine
        lat, old = sch.step(npred, old, t, sch.timesteps[i - 1] if i > 0 else None, lat, generator=g1, return_dict=False)
        lat = lat.to(torch.bfloat16)
        lat_ref, old_ref = orc.step(npred, old_ref, t, orc.timesteps[i - 1] if i > 0 else None, lat_ref,
                                    lambda shape, dtype: torch.randn(shape, generator=g2, device="cuda", dtype=dtype))
        lat_ref = lat_ref.to(torch.bfloat16)
        assert torch.equal(lat, lat_ref) and torch.equal(old, old_ref), i
    with pytest.raises(NotImplementedError):
        sch.step(npred, None, 999, None, lat, eta=0.5)
    with pytest.raises(NotImplementedError):
        sch.step(npred, None, 999, None, lat.float())


@pytest.mark.parametrize("zero2cond,dynamic,steps", [(False, False, 5), (True, True, 6)])
def test_denoise_loop_graph_equals_eager_equals_pipeline_oracle(built, zero2cond, dynamic, steps):
    """5 / 6 steps of the whole loop on the tiny config: one replayed CUDA graph per step == eager kernels == the
    pipeline's torch loop (pipeline_bindyouravatar.py:893-945) around the same CUDA transformer, on the same draws."""
    import bya_b200  # noqa: F401
    from bya_b200.denoise import DenoiseLoop
    from bya_b200.scheduler import CogVideoXDPMScheduler
    from bya_b200.synth import CONFIGS, make_inputs
    from oracle.dpm_oracle import DPMSchedulerOracle, denoise_loop_oracle
    from test_gpu_step import build

    cfg = dataclasses.replace(CONFIGS["c1"], batch=2)
    m = build(cfg)
    inp = make_inputs(cfg, 77, device="cuda", dtype=torch.bfloat16)
    inp.pop("timestep")
    hs = inp.pop("hidden_states")
    lat, img, bg = (hs[:1, :, 16 * k: 16 * (k + 1)].contiguous() for k in range(3))
    guidance = 6.0
    common = dict(prompt_embeds=inp["encoder_hidden_states"], image_rotary_emb=inp["image_rotary_emb"],
                  id_cond=inp["id_cond"], id_vit_hidden=inp["id_vit_hidden"], audio_embeds=inp["audio_embeds"],
                  af_matrix=inp["af_matrix"], num_inference_steps=steps)
    mk = lambda graph: DenoiseLoop(m, CogVideoXDPMScheduler.cogvideox_5b(), guidance_scale=guidance, use_dynamic_cfg=dynamic,
                                   zero2cond_cfg_flag=zero2cond, cuda_graph=graph)
    # eager kernels, with a per-step trace, drawing from a seeded device generator
    trace = []
    eager = mk(False)
    out_e = eager.run(lat, img, bg, generator=torch.Generator(device="cuda").manual_seed(3), trace=trace, **common).clone()
    noise = eager.state["noise"]
    assert len(trace) == steps and out_e.shape == lat.shape and out_e.dtype == torch.bfloat16
    # one CUDA graph replayed per step, same seed
    out_g = mk(True).run(lat, img, bg, generator=torch.Generator(device="cuda").manual_seed(3), **common).clone()
    assert torch.equal(out_g, out_e)
    # the pipeline's loop in torch around the same transformer, fed the same draws in the same order
    order = []
    coef = eager.state["coef"].cpu()
    for i in range(steps):
        order.append(noise[i, 0])
        if coef[i, 8] != 0:
            order.append(noise[i, 1])
    it = iter(order)

    def transformer(x, t, i):
        return m(hidden_states=x, timestep=t, denoise_step=i, return_dict=False,
                 encoder_hidden_states=inp["encoder_hidden_states"], image_rotary_emb=inp["image_rotary_emb"],
                 id_cond=inp["id_cond"], id_vit_hidden=inp["id_vit_hidden"], audio_embeds=inp["audio_embeds"],
                 af_matrix=inp["af_matrix"])[0].clone()

    otrace = []
    out_o = denoise_loop_oracle(transformer, DPMSchedulerOracle(), lat, img, bg, steps, guidance,
                                lambda shape, dtype: next(it).reshape(shape), use_dynamic_cfg=dynamic,
                                zero2cond_cfg_flag=zero2cond, trace=otrace)
    for i in range(steps):
        assert torch.equal(trace[i][0], otrace[i][0]), f"latents differ after step {i}"
        assert torch.equal(trace[i][1], otrace[i][1]), f"pred_original_sample differs after step {i}"
    assert torch.equal(out_o, out_e)
    assert next(it, None) is None and torch.isfinite(out_e.float()).all()


def test_denoise_loop_batches_the_conditions_like_the_pipeline(built):
    """conditions_cfg_batched=False == passing what pipeline_bindyouravatar.py:877-884 would have built."""
    import bya_b200  # noqa: F401
    from bya_b200.conditions import prepare_cfg_conditions
    from bya_b200.denoise import DenoiseLoop
    from bya_b200.scheduler import CogVideoXDPMScheduler
    from bya_b200.synth import CONFIGS, make_inputs
    from test_gpu_step import build

    cfg = CONFIGS["c1"]
    m = build(cfg)
    inp = make_inputs(cfg, 5, device="cuda", dtype=torch.bfloat16)
    hs = inp["hidden_states"]
    lat, img, bg = (hs[:1, :, 16 * k: 16 * (k + 1)].contiguous() for k in range(3))
    prompt = torch.cat([torch.zeros_like(inp["encoder_hidden_states"]), inp["encoder_hidden_states"]])
    outs = []
    for batched in (True, False):
        conds = (inp["id_cond"], inp["id_vit_hidden"], inp["audio_embeds"], inp["af_matrix"])
        if batched:
            conds = prepare_cfg_conditions(*conds, True, True)
        loop = DenoiseLoop(m, CogVideoXDPMScheduler.cogvideox_5b(), guidance_scale=4.0, zero2cond_cfg_flag=True)
        outs.append(loop.run(lat, img, bg, prompt, inp["image_rotary_emb"], *conds, num_inference_steps=3,
                             generator=torch.Generator(device="cuda").manual_seed(9),
                             conditions_cfg_batched=batched).clone())
    assert torch.equal(outs[0], outs[1]) and torch.isfinite(outs[0].float()).all()
    assert not torch.equal(outs[0], lat)
