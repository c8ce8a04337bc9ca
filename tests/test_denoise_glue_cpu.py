"""CPU (-m "not gpu"): the host half of SURVEY.md §8f row N1 — the CogVideoXDPMScheduler mirror and the loop oracle.

diffusers is not in this image and the reference holds no scheduler vectors, so the schedule is pinned by what the
published solver guarantees analytically (DPM-Solver++ SDE: the update preserves the mean sqrt(a) x0 and the variance
1 - a of the forward process; zero terminal SNR; trailing timesteps; x0-exact for a perfect predictor), and the
product's host table is compared with the oracle's independent restatement."""
import math

import pytest
import torch


@pytest.fixture(scope="module")
def sched():
    import bya_b200  # noqa: F401
    from bya_b200.scheduler import CogVideoXDPMScheduler

    return CogVideoXDPMScheduler.cogvideox_5b   # the CogVideoX-5B scheduler_config.json values (ctor defaults are diffusers')


def test_trailing_timesteps_and_zero_terminal_snr(sched):
    s = sched()
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(999, 0, -20))
    assert s.timesteps.dtype == torch.int64
    assert float(s.alphas_cumprod[-1]) == 0.0                       # zero terminal SNR
    assert s.alphas_cumprod.dtype == torch.float64
    assert abs(float(s.alphas_cumprod[0]) - (1 - 0.00085)) < 1e-12  # the first entry is a fixed point of the rescale
    assert bool((s.alphas_cumprod[1:] < s.alphas_cumprod[:-1]).all())
    s.set_timesteps(7)
    assert s.timesteps.tolist() == [999, 856, 713, 570, 428, 285, 142]
    assert s.scale_model_input("x", 3) == "x" and s.order == 1 and s.init_noise_sigma == 1.0
    with pytest.raises(ValueError):
        s.set_timesteps(1001)


def test_from_config_drops_variance_type_like_infer_py(sched, tmp_path):
    import json

    from bya_b200.scheduler import COGVIDEOX_5B_SCHEDULER_CONFIG, CogVideoXDPMScheduler

    s = CogVideoXDPMScheduler.from_config(dict(sched().config), variance_type="fixed_small", snr_shift_scale=3.0)
    assert s.config.snr_shift_scale == 3.0 and "variance_type" not in s.config
    with pytest.raises(ValueError):
        sched(prediction_type="flow")
    # bare constructor = the published class's defaults, not the 5B checkpoint's values
    d = CogVideoXDPMScheduler().config
    assert (d.prediction_type, d.timestep_spacing, d.rescale_betas_zero_snr, d.snr_shift_scale) == ("epsilon", "leading", False, 3.0)
    # infer.py:202: CogVideoXDPMScheduler.from_pretrained(model_path, subfolder="scheduler")
    (tmp_path / "scheduler").mkdir()
    cfg = dict(COGVIDEOX_5B_SCHEDULER_CONFIG, _class_name="CogVideoXDPMScheduler", _diffusers_version="0.30.0.dev0")
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(cfg))
    p = CogVideoXDPMScheduler.from_pretrained(str(tmp_path), subfolder="scheduler")
    assert dict(p.config) == dict(sched().config)
    with pytest.raises(OSError):
        CogVideoXDPMScheduler.from_pretrained(str(tmp_path), subfolder="nope")


@pytest.mark.parametrize("steps,shift", [(50, 1.0), (50, 3.0), (12, 1.0), (1000, 1.0)])
def test_solver_identities(sched, steps, shift):
    """mean:  mult0 sqrt(a_t) - mult1 = sqrt(a_prev);  variance:  mult0^2 (1 - a_t) + mult_noise^2 = 1 - a_prev;
    second order:  mult2 - mult3 = 1 (the extrapolation keeps a constant prediction)."""
    s = sched(snr_shift_scale=shift)
    s.set_timesteps(steps)
    ts = s.timesteps.tolist()
    stride = 1000 // steps
    for i, t in enumerate(ts):
        row = s.step_coefficients(t, ts[i - 1] if i else None, i > 0, 6.0)
        g, sa, sb, m0, m1, m2, m3, mn, second = row[:9]
        a_t = float(s.alphas_cumprod[t])
        a_p = float(s.alphas_cumprod[t - stride]) if t - stride >= 0 else 1.0
        assert g == 6.0 and abs(sa - math.sqrt(a_t)) < 1e-6 and abs(sb - math.sqrt(1 - a_t)) < 1e-6
        assert abs(m0 * math.sqrt(a_t) - m1 - math.sqrt(a_p)) < 2e-6, (i, t)
        assert abs(m0 * m0 * (1 - a_t) + mn * mn - (1 - a_p)) < 2e-6, (i, t)
        assert second == float(i > 0 and t - stride >= 0)
        if second:
            assert abs(m2 - m3 - 1.0) < 1e-6 and m3 >= 0
        else:
            assert m2 == 0.0 and m3 == 0.0
        assert all(math.isfinite(v) for v in row[:9]) and (math.isfinite(row[9]) or a_t == 0.0)
        assert a_t == 0.0 or abs(row[9] * sa - 1.0) < 1e-6
    # last step of a trailing schedule lands on alpha = 1: x_0 = pred, no noise
    assert ts[-1] - stride < 0 and (m0, m1, mn) == (0.0, -1.0, 0.0)


def test_host_table_equals_oracle_restatement(sched):
    from oracle.dpm_oracle import DPMSchedulerOracle

    for kw in (dict(), dict(snr_shift_scale=3.0), dict(beta_schedule="linear", rescale_betas_zero_snr=False),
               dict(timestep_spacing="leading"), dict(timestep_spacing="linspace", set_alpha_to_one=False)):
        s, o = sched(**kw), DPMSchedulerOracle(**kw)
        assert torch.equal(s.alphas_cumprod, o.alphas_cumprod)
        s.set_timesteps(25)
        o.set_timesteps(25)
        assert torch.equal(s.timesteps, o.timesteps)
        ts = o.timesteps.tolist()
        for i, t in enumerate(ts):
            a_t, mult, mult_noise, prev_t = o.coefficients(t, ts[i - 1] if i else None)
            row = s.step_coefficients(t, ts[i - 1] if i else None, i > 0)
            want = [a_t**0.5, (1 - a_t) ** 0.5, mult[0], mult[1], mult_noise]
            got = [row[1], row[2], row[3], row[4], row[7]]
            for w, g in zip(want, got):
                assert float(w.to(torch.float32)) == g
            if i and prev_t >= 0:
                assert float(mult[2].to(torch.float32)) == row[5] and float(mult[3].to(torch.float32)) == row[6]


def test_loop_oracle_recovers_x0_with_a_perfect_v_predictor():
    """With v = sqrt(a) eps - sqrt(1 - a) x0 for the eps that explains the current latents, pred_original_sample is x0
    on every step, the second-order extrapolation keeps it, and the last step returns it — whatever noise is drawn."""
    from oracle.dpm_oracle import DPMSchedulerOracle, denoise_loop_oracle

    torch.manual_seed(0)
    o = DPMSchedulerOracle()
    x0 = torch.randn(1, 3, 16, 4, 6, dtype=torch.float64)
    img = torch.zeros(1, 3, 16, 4, 6, dtype=torch.float64)

    def model(x, t, i):
        a = o.alphas_cumprod[int(t[0])]
        lat = x[:, :, :16]
        assert x.shape == (2, 3, 48, 4, 6) and t.shape == (2,)
        if float(a) == 0.0:
            return -x0.expand(2, -1, -1, -1, -1).clone()    # pure noise: v = -x0 gives pred = x0
        eps = (lat - a**0.5 * x0) / (1 - a) ** 0.5
        return (a**0.5 * eps - (1 - a) ** 0.5 * x0)

    draws = []

    def randn(shape, dtype):
        draws.append(1)
        return torch.randn(shape, dtype=dtype)

    # float64 end to end (the oracle's .float() on the model output is the only fp32 rounding)
    out = denoise_loop_oracle(model, o, torch.randn(1, 3, 16, 4, 6, dtype=torch.float64), img, img, 10, 6.0, randn,
                              out_dtype=torch.float64)
    assert float((out - x0).abs().max()) < 1e-5
    assert len(draws) == 1 + 2 * 8 + 1   # first and last step draw once, the second-order steps twice


def test_dynamic_guidance_matches_pipeline_formula():
    import bya_b200  # noqa: F401
    from bya_b200.denoise import DenoiseLoop
    from bya_b200.scheduler import CogVideoXDPMScheduler
    from oracle.dpm_oracle import dynamic_guidance

    s = CogVideoXDPMScheduler.cogvideox_5b()
    s.set_timesteps(50)
    loop = DenoiseLoop(None, s, guidance_scale=6.0, use_dynamic_cfg=True)
    tab = loop.coefficient_table(50)
    assert tab.shape == (50, 12) and tab.dtype == torch.float32
    for i, t in enumerate(s.timesteps.tolist()):
        assert float(tab[i, 0]) == float(torch.tensor(dynamic_guidance(6.0, 50, t), dtype=torch.float32))
    assert float(DenoiseLoop(None, s, guidance_scale=6.5).coefficient_table(50)[7, 0]) == 6.5
    assert tab[0, 8] == 0 and tab[-1, 8] == 0 and bool((tab[1:-1, 8] == 1).all())


def test_up_front_noise_draws_follow_the_reference_order():
    """`DenoiseLoop.draw_noise` takes every step's randn up front; with the same seed it must hand step i exactly the
    draws the step-by-step loop (oracle: pipeline :934-943 + scheduler.step) would have taken at step i — one on the
    first / last step, two on second-order steps (the second one is the one that is used)."""
    import bya_b200  # noqa: F401
    from bya_b200.denoise import DenoiseLoop
    from bya_b200.scheduler import CogVideoXDPMScheduler, randn_tensor
    from oracle.dpm_oracle import DPMSchedulerOracle, denoise_loop_oracle

    shape, steps = (1, 2, 16, 4, 6), 7
    sch = CogVideoXDPMScheduler.cogvideox_5b()
    sch.set_timesteps(steps)
    loop = DenoiseLoop(None, sch, guidance_scale=3.0)
    noise = loop.draw_noise(shape, loop.coefficient_table(steps), torch.Generator().manual_seed(21), "cpu")
    assert noise.shape == (steps, 2, 2 * 16 * 4 * 6) and noise.dtype == torch.bfloat16

    g2, per_step, current = torch.Generator().manual_seed(21), [], []

    def randn(s, dtype):
        current.append(randn_tensor(s, g2, "cpu", dtype))
        return current[-1]

    def model(x, t, i):
        if current:
            per_step.append(list(current))
            current.clear()
        return torch.zeros(2, 2, 16, 4, 6, dtype=torch.bfloat16)

    lat = torch.zeros(shape, dtype=torch.bfloat16)
    denoise_loop_oracle(model, DPMSchedulerOracle(), lat, lat, lat, steps, 3.0, randn)
    per_step.append(list(current))
    # 7 trailing steps end on t = 142 with prev_timestep = 0: still a second-order step (unlike 50 steps, which end at -1)
    assert [len(d) for d in per_step] == [1] + [2] * (steps - 1)
    assert [len(d) for d in per_step] == [1 + int(v) for v in loop.coefficient_table(steps)[:, 8]]
    for i, draws in enumerate(per_step):
        for slot, d in enumerate(draws):
            assert torch.equal(noise[i, slot], d.reshape(-1)), (i, slot)
        if len(draws) == 1:
            assert float(noise[i, 1].abs().max()) == 0.0
