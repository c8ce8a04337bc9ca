"""The oracle is only as good as its pins: restated torch oracle vs the reference-produced goldens (and vs the
verbatim reference when /root/reference is mounted); C mask oracle vs torch and vs the reference-produced goldens."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def c1_model_state():
    import bya_b200  # noqa: F401
    from bya_b200.synth import CONFIGS, fill_module
    from bya_b200.transformer import BindyouravatarTransformer3DModel

    cfg = CONFIGS["c1"]
    m = BindyouravatarTransformer3DModel(**cfg.ctor_kwargs()).eval()
    m.router.set_grid(cfg.frames, cfg.grid_h, cfg.grid_w)
    fill_module(m, 0)
    return cfg, {k: v for k, v in m.state_dict().items()}


def test_restated_oracle_matches_reference_golden_config1(c1_model_state):
    """tests/golden/step_c1_soft.pt was produced by the UNMODIFIED reference forward (oracle/make_goldens.py)."""
    from bya_b200.synth import make_inputs
    from oracle import restated

    cfg, sd = c1_model_state
    g = torch.load(os.path.join(GOLD, "step_c1_soft.pt"))
    taps = {}
    out = restated.step(sd, cfg, **make_inputs(cfg, g["input_seed"]), taps=taps)
    assert float((out - g["output"]).abs().max()) < 2e-4
    assert float((taps["ca0.router"] - g["router"]).abs().max()) < 1e-5
    assert float((taps["block0.video"][:, ::39] - g["block0.video"]).abs().max()) < 1e-4
    assert float((taps["face_tokens"][:, :, ::4] - g["face_tokens"]).abs().max()) < 1e-4


def _cpu_model_state(cfg):
    import bya_b200  # noqa: F401
    from bya_b200.synth import fill_module
    from bya_b200.transformer import BindyouravatarTransformer3DModel

    m = BindyouravatarTransformer3DModel(**cfg.ctor_kwargs()).eval()
    m.router.set_grid(cfg.frames, cfg.grid_h, cfg.grid_w)
    fill_module(m, 0)
    return m, dict(m.state_dict())


def test_restated_oracle_matches_reference_golden_L4_interval2_batch2():
    """tests/golden/step_L4_int2_b2.pt (UNMODIFIED reference, round 2): 4 blocks, cross_attn_interval=2 — layers 1 and 3
    take the audio weights from the routing of the PREVIOUS cross-attention layer (transformer.py:858-863) — CFG batch 2
    with the unconditional branch's audio zeroed.  Pins the layer loop, not just one block."""
    import dataclasses

    from bya_b200.synth import CONFIGS, make_inputs
    from oracle import restated

    cfg = dataclasses.replace(CONFIGS["c1"], num_layers=4, cross_attn_interval=2, batch=2)
    _, sd = _cpu_model_state(cfg)
    g = torch.load(os.path.join(GOLD, "step_L4_int2_b2.pt"))
    inp = make_inputs(cfg, g["input_seed"])
    inp["audio_embeds"][0] = 0
    taps = {}
    out = restated.step(sd, cfg, **inp, taps=taps)
    assert float((out - g["output"]).abs().max()) < 5e-4
    for i in range(4):
        assert float((taps[f"block{i}.video"][:, ::13, ::7] - g[f"block{i}.video"]).abs().max()) < 2e-4, i
    # router calls of the reference: [ca0 b0, ca0 b1, ca1 b0, ca1 b1]; the oracle taps batch element 0
    assert float((taps["ca0.router"] - g["router"][0]).abs().max()) < 1e-5
    assert float((taps["ca1.router"] - g["router"][2]).abs().max()) < 1e-5
    assert "ca2.router" not in taps


@pytest.mark.parametrize("name,kw", [("step_c1_learnedpos.pt", dict(use_learned_positional_embeddings=True)),
                                     ("step_c1_sincos.pt", dict(use_rotary_positional_embeddings=False))])
def test_positional_embedding_configurations_vs_reference_golden(name, kw):
    """The two other constructions of the reference's patch embedding (models/transformer.py:370-392): the learned
    `patch_embed.pos_embedding` checkpoint buffer of the CogVideoX-5B-I2V lineage (with RoPE), and the analytic sincos
    table of a model built without RoPE.  Goldens by the UNMODIFIED reference on the diffusers shim; the product-side
    table (`bya_b200.modules.sincos_positional_embedding`) must equal the shim's."""
    import dataclasses

    from bya_b200.synth import CONFIGS, make_inputs
    from oracle import restated

    cfg = dataclasses.replace(CONFIGS["c1"], **kw)
    m, sd = _cpu_model_state(cfg)
    assert ("patch_embed.pos_embedding" in sd) == bool(kw.get("use_learned_positional_embeddings"))
    g = torch.load(os.path.join(GOLD, name))
    assert float((m.patch_embed.pos_embedding[:, ::61, ::17] - g["pos_embedding_sub"]).abs().max()) < 1e-6
    inp = make_inputs(cfg, g["input_seed"])
    assert (inp["image_rotary_emb"] is None) == (not cfg.use_rotary_positional_embeddings)
    out = restated.step(sd, cfg, **inp)
    assert float((out - g["output"]).abs().max()) < 2e-4


def test_restated_oracle_generalises_consistently(c1_model_state):
    """C>2 audio-weight rule reduces to the reference's swap at C=2; frame-OR matches transformer.py:815-818."""
    from oracle import restated

    r = torch.rand(50, 2)
    af = torch.eye(2)
    w = restated.audio_weights(af, r)
    assert torch.equal(w, 1 - r[:, [1, 0]])
    w3 = restated.audio_weights(torch.eye(3), torch.rand(20, 3))
    assert w3.shape == (20, 3)
    lg = (torch.rand(1, 4 * 3 * 5, 2) > 0.7).float()
    o = restated.frame_or(lg, 4, 3, 5).view(4, 15, 2)
    assert torch.equal(o[0], lg.view(4, 15, 2).max(0).values) and torch.equal(o[0], o[3])


@pytest.mark.reference
def test_restated_oracle_matches_verbatim_reference(c1_model_state):
    from oracle.reference_harness import build_reference_model, reference_available, run_reference

    if not reference_available():
        pytest.skip("reference tree not mounted (GPU box)")
    from bya_b200.synth import make_inputs
    from oracle import restated

    cfg, sd = c1_model_state
    m = build_reference_model(cfg, seed=0)
    assert set(m.state_dict()) == set(sd)
    inp = make_inputs(cfg, 77)
    ref = run_reference(m, inp)
    out = restated.step(sd, cfg, **inp)
    assert float((out - ref).abs().max()) < 2e-5


# ------------------------------------------------------------------------------------------------ mask oracle
@pytest.fixture(scope="module")
def mask_lib(built):
    return ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libmask_oracle.so"))


def _c_trilinear(lib, x, F_, gh, gw):
    T, H, W = x.shape
    y = np.zeros((F_, gh, gw), np.float32)
    xn = np.ascontiguousarray(x.numpy(), np.float32)
    lib.bya_oracle_trilinear(xn.ctypes.data_as(ctypes.c_void_p), T, H, W, y.ctypes.data_as(ctypes.c_void_p), F_, gh, gw)
    return y


def test_c_trilinear_is_float_exact_vs_torch(mask_lib):
    torch.manual_seed(0)
    for (T, H, W, F_, gh, gw) in [(49, 96, 144, 13, 6, 9), (97, 100, 150, 25, 7, 11), (49, 64, 80, 13, 30, 45),
                                  (10, 33, 47, 13, 8, 12), (13, 30, 45, 13, 30, 45), (5, 8, 8, 1, 1, 1)]:
        x = torch.rand(T, H, W)
        ref = F.interpolate(x[None, None], size=(F_, gh, gw), mode="trilinear", align_corners=False)[0, 0].numpy()
        assert np.array_equal(ref, _c_trilinear(mask_lib, x, F_, gh, gw)), (T, H, W, F_, gh, gw)


def c_masks_to_routing(lib, masks, F_, gh, gw, frame_or=False):
    C, T, H, W = masks.shape
    n = F_ * gh * gw
    idx = np.zeros(n, np.int64)
    lg = np.zeros((n, C), np.float32)
    m = np.ascontiguousarray(masks, np.uint8)
    lib.bya_oracle_masks_to_routing(m.ctypes.data_as(ctypes.c_void_p), C, T, H, W, F_, gh, gw,
                                    idx.ctypes.data_as(ctypes.c_void_p), lg.ctypes.data_as(ctypes.c_void_p), int(frame_or))
    return idx, lg


@pytest.mark.parametrize("kind", ["moving", "static", "overlap", "speckle"])
def test_c_mask_oracle_bit_exact_vs_reference_golden(mask_lib, kind):
    """Goldens = the reference's own process_masks_to_routing_logits (util/utils.py:871-936) on PNG dirs."""
    import bya_b200  # noqa: F401
    from bya_b200.synth import tracking_masks

    g = torch.load(os.path.join(GOLD, f"masks_{kind}.pt"))
    masks = tracking_masks(kind)
    assert int(masks.astype(np.int64).sum()) == g["mask_checksum"]
    idx, lg = c_masks_to_routing(mask_lib, masks, 13, 30, 45)
    assert np.array_equal(lg, g["routing_logits"][0].numpy().astype(np.float32))
    assert set(np.unique(idx)) <= {-1, 0, 1}
    if kind == "moving":  # exact-0.5 ties exist in this case and must resolve to "not inside" (strict >)
        assert int((lg.sum(1) == 0).sum()) > 0


def test_torch_restatement_of_mask_path_agrees(mask_lib):
    import bya_b200  # noqa: F401
    from bya_b200.synth import tracking_masks
    from oracle import restated

    masks = tracking_masks("overlap", T=17, H=64, W=96)
    idx_t, lg_t = restated.routing_from_masks(torch.from_numpy(masks), 5, 8, 12)
    idx_c, lg_c = c_masks_to_routing(mask_lib, masks, 5, 8, 12)
    assert np.array_equal(idx_t[0].numpy(), idx_c) and np.array_equal(lg_t[0].numpy(), lg_c)
    _, lg_or = c_masks_to_routing(mask_lib, masks, 5, 8, 12, frame_or=True)
    assert np.array_equal(restated.frame_or(lg_t, 5, 8, 12)[0].numpy(), lg_or)
