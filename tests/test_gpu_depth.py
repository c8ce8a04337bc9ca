"""-m gpu: FULL-DEPTH parity of the headline configuration: 42 layers, 13x30x45 grid (17 776 tokens),
cross_attn_interval 2, compared layer by layer with the fp32 oracle run on the same GPU over the same bf16-rounded
weights — next to the torch-bf16 evaluation of the same oracle (the noise floor of the format).  The per-layer table of
each run is written to gpurun_out/depth_parity_*.json (summarised under profiles/).

Tolerance (BASELINE.json north_star): cosine >= 0.999 on the noise prediction; max-abs stated below; the routing
outputs of the hard-mask path are bit-exact (tests/test_gpu_kernels.py, tests/test_gpu_step.py)."""
import dataclasses
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dump(name, res):
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "w") as f:
            json.dump(res, f, indent=1)
    except OSError:
        pass


def _check(res, min_cos, max_rel):
    o = res["output"]
    assert o["finite"]
    worst = min(res["taps"].items(), key=lambda kv: kv[1]["cos"])
    print(f"output: cos {o['cos']:.6f} (torch-bf16 {o['cos_bf16']:.6f})  rel max-abs {o['rel_max']:.4f} "
          f"(torch-bf16 {o['rel_max_bf16']:.4f});  worst tap {worst[0]}: {worst[1]}")
    assert o["cos"] >= min_cos, o
    assert o["rel_max"] <= max_rel, o
    for k in ("face_tokens", "audio_ctx"):      # the per-generation prologue (own kernels)
        assert res["taps"][k]["cos"] >= 0.9995 and res["taps"][k]["rel_max"] <= 0.03, (k, res["taps"][k])
    for k, r in res["taps"].items():
        if k.endswith(".video"):
            assert r["cos"] >= min_cos, (k, r)
        if k.endswith(".router"):   # soft routing in (0,1): absolute error (SURVEY.md §0.6 — bf16 tolerance, not bit-exact)
            assert r["rel_max"] <= 0.06, (k, r)
    # never further from the truth than 3x the distance of torch's own bf16 evaluation (+ slack for tiny distances)
    assert 1 - o["cos"] <= 3 * (1 - o["cos_bf16"]) + 2e-4, o


def test_c2_full_depth_soft_router_vs_fp32_oracle(built):
    """configs[1]: 42 layers, 49 frames 480x720, 2 characters, B=1, learned soft router, face + audio cross-attention."""
    from bya_b200.synth import CONFIGS
    from depth_parity import depth_parity

    res = depth_parity(CONFIGS["c2"])
    _dump("depth_parity_c2_soft.json", res)
    assert len([k for k in res["taps"] if k.startswith("block")]) == 42
    assert len([k for k in res["taps"] if k.endswith(".router")]) == 21
    _check(res, 0.999, 0.08)


def test_c3_full_depth_forced_masks_cfg_batch2(built):
    """configs[2] geometry: CFG batch 2 (unconditional branch with zeroed audio), stage-2 forced hard masks (router
    skipped, frame-OR, bit-exact audio weights), 42 layers at the full grid."""
    from bya_b200.synth import CONFIGS
    from depth_parity import depth_parity

    res = depth_parity(CONFIGS["c3"], forced_masks=True)
    _dump("depth_parity_c3_forced.json", res)
    _check(res, 0.999, 0.08)
